#!/bin/bash
# round 2, GPU call D: staged four-block calls in double precision (P+7 delay-line slots), branch-free dither walker
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
python -m pytest tests -m gpu -q -x > gpurun_out/r02d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02d_pytest.log
tail -4 gpurun_out/r02d_pytest.log
python - <<'PY'
import importlib, json, sys
sys.path.insert(0, "tools")
import torch
import bench_configs as bc
pkg = importlib.import_module("foo-dsp-bfir_b200")
for s in (1, 64):
    print(json.dumps(bc.dither_timing(pkg, torch, streams=s, blocks=100)))
PY
( time python bench.py --steps 100 --warmup 4 --no-configs ) > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; echo "bench exit $?"
tail -4 gpurun_out/r02d_bench.err
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r02d_bench.json").read().strip().splitlines()[-1])
r = j["roofline"]
print("value", j["value"], "ms/step", j["ms_per_step"], "e2e", j["e2e"]["value"], "frac", r["frac"], "mac ms", r["avg_launch_ms"])
print("serial", r["serial_pass"], "share", r["step_share"])
print("pairs", r["two_blocks_per_launch"]["staged"], r["two_blocks_per_launch"]["serial"]["value"])
print("check", j["output_check"], "launches", j["gpu_launches"])
PY
python tools/pair_ncu_target.py --quads > gpurun_out/r02d_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:partition_mac_multi -s 2 -c 2 -f -o gpurun_out/r02d_mac_quad python tools/pair_ncu_target.py --quads > gpurun_out/r02d_ncu.log 2>&1
tail -2 gpurun_out/r02d_ncu.log
