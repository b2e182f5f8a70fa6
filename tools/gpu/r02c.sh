#!/bin/bash
# round 2, GPU call C: dither kernel rewrite + double quad partition sum: tests, kernel times
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "dither or quad or cbuf2raw or integer or S16 or stats" > gpurun_out/r02c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c_pytest.log
tail -4 gpurun_out/r02c_pytest.log
rm -f gpurun_out/r02c_kt.jsonl
for sp in 1 2 4; do
  BFIR_MAC_SPLIT=$sp python tools/kernel_times.py --tag cfg1x16_quads_split$sp --steps 200 --quads >> gpurun_out/r02c_kt.jsonl 2>&1
done
python tools/kernel_times.py --tag cfg1x16_quads --steps 200 --quads >> gpurun_out/r02c_kt.jsonl 2>&1
python tools/kernel_times.py --tag cfg1x16_pairs --steps 200 --pairs >> gpurun_out/r02c_kt.jsonl 2>&1
cat gpurun_out/r02c_kt.jsonl
python - <<'PY'
import importlib, json, sys
sys.path.insert(0, "tools")
import torch
import bench_configs as bc
pkg = importlib.import_module("foo-dsp-bfir_b200")
for s in (1, 64, 4096):
    print(json.dumps(bc.dither_timing(pkg, torch, streams=s, blocks=100)))
PY
