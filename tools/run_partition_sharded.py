#!/usr/bin/env python
"""Partition-sharded long-filter convolution across the GPUs of one box (BASELINE configs[4] geometry):
every rank convolves P/world partitions of all filters, the partial spectra are summed with ONE NCCL
all-reduce per block, every rank runs the (cheap) output stage. Launch with torchrun:

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/run_partition_sharded.py \
        --L 32768 --P 64 --size 32 --blocks 20 [--check]

--check compares against an unsharded engine on rank 0 (same inputs). Prints one JSON line (rank 0)."""
import argparse
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--L", type=int, default=32768)
    ap.add_argument("--P", type=int, default=64)
    ap.add_argument("--size", dest="n", type=int, default=32, help="inputs = filters = outputs")
    ap.add_argument("--blocks", type=int, default=20)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--fused", action="store_true", help="peer-store fused reduce instead of the NCCL all-reduce")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = importlib.import_module("foo-dsp-bfir_b200")
    sh = importlib.import_module("foo-dsp-bfir_b200.sharding")
    L, P, n = a.L, a.P, a.n
    rng = np.random.default_rng(5)
    taps = L * P
    env = np.exp(-6.9 * np.arange(taps) / taps).astype(np.float32)
    base = rng.standard_normal(taps).astype(np.float32) * env
    base /= np.sqrt(np.sum(base.astype(np.float64) ** 2))
    h = [np.roll(base, 17 * f) * np.float32(1.0 + 0.01 * f) for f in range(n)]
    gin = rng.standard_normal((n, n)) / np.sqrt(n)
    gout = rng.standard_normal((n, n)) / np.sqrt(n)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    drv, eng = sh.make_partition_sharded(pkg, L, P, 4, n, pkg.FLOAT_LE, pkg.FLOAT_LE, 48000, False, h, device=local,
                                         xbar_inputs=n, xbar_outputs=n, in_gains=gin, out_gains=gout)
    own_first, own_count = 0, n
    if a.fused:
        assert world > 1, "--fused needs at least two ranks"
        drv = sh.FusedPartitionShardedEngine(eng)
        own_first, own_count = drv.own_first, drv.own_count
    full = None
    if a.check and rank == 0:
        full = pkg.Brutefir(L, P, 4, n, pkg.FLOAT_LE, pkg.FLOAT_LE, 48000, False, device=local, n_groups=1, xbar_inputs=n, xbar_outputs=n)
        assert full.set_coeff(h, P) == 0
        full.set_crossbar(gin, gout)
        full.set_stream(stream.cuda_stream)
    d_in = [torch.from_numpy(np.random.default_rng(0xB200 + b).uniform(-1, 1, L * n).astype(np.float32)).cuda() for b in range(4)]
    d_out = torch.empty(L * own_count, dtype=torch.float32, device="cuda")
    d_ref = torch.empty(L * n, dtype=torch.float32, device="cuda")
    worst = 0.0
    warm = P          # every timed block convolves all partitions
    for b in range(warm):
        drv.run_device(d_in[b % 4], d_out)
        if full is not None:
            full.run_device(d_in[b % 4], d_ref)
    assert drv.sync() == 0
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for b in range(a.blocks):
        drv.run_device(d_in[(warm + b) % 4], d_out)
    ev1.record(stream)
    assert drv.sync() == 0
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / a.blocks
    if full is not None:
        for b in range(a.blocks):
            full.run_device(d_in[(warm + b) % 4], d_ref)
        assert full.sync() == 0
        torch.cuda.synchronize()
        y, r = d_out.double().reshape(L, own_count), d_ref.double().reshape(L, n)[:, own_first:own_first + own_count]
        worst = float(torch.sqrt(torch.mean((y - r) ** 2) / torch.mean(r ** 2)))
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        part = (P + world - 1) // world
        print(json.dumps({"workload": "cfg4-geometry: %dx%d crossbar, L %d, P %d (%d taps), float" % (n, n, L, P, taps),
                          "world": world, "reduce": "fused peer stores + 1-element barrier" if a.fused else "NCCL all-reduce", "partitions_per_rank": part, "ms_per_block": ms,
                          "Msamples_s": n * L / (ms * 1e-3) / 1e6, "reduce_bytes": n * 2 * L * 4,
                          "rel_rms_vs_unsharded": worst if a.check else None,
                          "prefill": "timed after a %d-block prefill (all partitions active)" % warm}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
