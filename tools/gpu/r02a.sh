#!/bin/bash
# round 2, GPU call A: full GPU suite with parity records, kernel times of the transforms, ncu summary of the
# cfg1 x 16 transform kernels (grid 128)
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_smi.csv 2>&1
python -m pytest tests -m gpu -x -q -s > gpurun_out/r02a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02a_pytest.log
tail -5 gpurun_out/r02a_pytest.log
for extra in "" "--pairs"; do
  python tools/kernel_times.py --tag cfg1x16$extra --steps 200 $extra >> gpurun_out/r02a_kernel_times.jsonl 2>&1
done
python tools/kernel_times.py --tag cfg3_1024 --channels 2 --realsize 4 --L 4096 --P 16 --streams 1024 --steps 100 >> gpurun_out/r02a_kernel_times.jsonl 2>&1
python tools/kernel_times.py --tag cfg1x1 --streams 1 --steps 200 >> gpurun_out/r02a_kernel_times.jsonl 2>&1
cat gpurun_out/r02a_kernel_times.jsonl
python tools/kernel_times.py --tag ncu --steps 8 > gpurun_out/r02a_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rfft_ -s 70 -c 4 -f -o gpurun_out/r02a_fft_f64_13 python tools/kernel_times.py --tag ncu --steps 8 > gpurun_out/r02a_ncu.log 2>&1
tail -3 gpurun_out/r02a_ncu.log
python bench.py --steps 100 --warmup 4 --no-latency > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench exit $?"
cat gpurun_out/r02a_bench.json | head -c 3000
