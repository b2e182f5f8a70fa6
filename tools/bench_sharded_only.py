#!/usr/bin/env python
"""Only the `partition_sharded` block of bench.py (tools/bench_configs.py: cfg4 with its partitions sharded over the
ranks), for tuning without the rest of the bench:

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/bench_sharded_only.py [--P 512] [--blocks 20]
Prints one JSON line (rank 0)."""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--L", type=int, default=32768)
    ap.add_argument("--P", type=int, default=512)
    ap.add_argument("--size", dest="n", type=int, default=32)
    ap.add_argument("--blocks", type=int, default=20)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    import bench_configs as bc
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = importlib.import_module("foo-dsp-bfir_b200")
    sh = importlib.import_module("foo-dsp-bfir_b200.sharding")
    pkg.load_library()
    out = bc.partition_sharded(pkg, sh, torch, dist, rank, world, local, blocks=a.blocks, L=a.L, P=a.P, n=a.n)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
