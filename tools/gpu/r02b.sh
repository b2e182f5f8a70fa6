#!/bin/bash
# round 2, GPU call B: full GPU suite (with the stream-ordering fixes), bench with the configs block, reference arm
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
python -m pytest tests -m gpu -q -s > gpurun_out/r02b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02b_pytest.log
tail -5 gpurun_out/r02b_pytest.log
( time python bench.py --steps 20 --warmup 3 ) > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench exit $?"
tail -4 gpurun_out/r02b_bench.err
( time python bench.py --impl reference --steps 20 --warmup 3 ) > gpurun_out/r02b_bench_ref.json 2> gpurun_out/r02b_bench_ref.err; echo "ref exit $?"
cat gpurun_out/r02b_bench_ref.json | head -c 1500
tail -4 gpurun_out/r02b_bench_ref.err
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r02b_bench.json").read().strip().splitlines()[-1])
print("value", j["value"], "e2e", j["e2e"]["value"], "frac", j["roofline"]["frac"])
print(json.dumps(j.get("configs"))[:3000])
print(json.dumps(j.get("partition_sharded"))[:1500])
PY
