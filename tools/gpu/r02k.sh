#!/bin/bash
# round 2, GPU call K (2 GPUs): four-block calls on partition shards (emulated ranks on one GPU, then two real ranks in bench.py)
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "fused or xbar or shard or quad" > gpurun_out/r02k_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02k_pytest.log
grep -E "^E|FAILED" gpurun_out/r02k_pytest.log | head -20
tail -3 gpurun_out/r02k_pytest.log
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 --no-latency --no-cpu-baseline ) > gpurun_out/r02k_bench_2gpu.json 2> gpurun_out/r02k_bench_2gpu.err; echo "bench exit $?"
tail -5 gpurun_out/r02k_bench_2gpu.err
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r02k_bench_2gpu.json").read().strip().splitlines()[-1])
print("value", j["value"], "e2e", j["e2e"]["value"])
print(json.dumps(j.get("partition_sharded"), indent=1)[:4000])
PY
