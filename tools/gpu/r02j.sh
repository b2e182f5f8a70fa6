#!/bin/bash
# round 2, GPU call J: double-precision 65536-point transforms (four CTAs, cluster + DSMEM)
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "32768 or largest or time2freq or coeffs2cbuf or crossfade" > gpurun_out/r02j_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02j_pytest.log
grep -E "^E|FAILED" gpurun_out/r02j_pytest.log | head -20
tail -4 gpurun_out/r02j_pytest.log
