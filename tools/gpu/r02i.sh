#!/bin/bash
# round 2, GPU call I: N3 (offline tools vs the reference's preprocessor.cpp) and N4 (coefficient-set routing), full suite
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
python -m pytest tests -m gpu -q -x -s > gpurun_out/r02i_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02i_pytest.log
grep -E "convolve_impulses|calculate_attenuation|preprocessor parity" gpurun_out/r02i_pytest.log
tail -4 gpurun_out/r02i_pytest.log
