#!/bin/bash
# round 2, GPU call E (2 GPUs): full GPU suite, then the 2-GPU bench with the configs block (cfg3 stream-sharded, cfg4 partition-sharded)
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
python -m pytest tests -m gpu -q -x > gpurun_out/r02e_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02e_pytest.log
tail -4 gpurun_out/r02e_pytest.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 ) > gpurun_out/r02e_bench_2gpu.json 2> gpurun_out/r02e_bench_2gpu.err; echo "bench exit $?"
tail -5 gpurun_out/r02e_bench_2gpu.err
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r02e_bench_2gpu.json").read().strip().splitlines()[-1])
print("value", j["value"], "e2e", j["e2e"]["value"], "frac", j["roofline"]["frac"])
print(json.dumps(j.get("configs"))[:2500])
print(json.dumps(j.get("partition_sharded"))[:2500])
PY
