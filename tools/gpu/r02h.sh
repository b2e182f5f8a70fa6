#!/bin/bash
# round 2, GPU call H: TMA bulk staging in the transforms + branch-free store statistics: suite, A/B, ncu of the transforms
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
python -m pytest tests -m gpu -q -x > gpurun_out/r02h_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02h_pytest.log
tail -4 gpurun_out/r02h_pytest.log
rm -f gpurun_out/r02h_kt.jsonl
for tma in 0 1; do
  BFIR_FFT_TMA=$tma python tools/kernel_times.py --tag cfg1x16_quads_tma$tma --steps 200 --quads >> gpurun_out/r02h_kt.jsonl 2>&1
  BFIR_FFT_TMA=$tma python tools/kernel_times.py --tag cfg3_1024_quads_tma$tma --channels 2 --realsize 4 --L 4096 --P 16 --streams 1024 --steps 100 --quads >> gpurun_out/r02h_kt.jsonl 2>&1
  BFIR_FFT_TMA=$tma python tools/kernel_times.py --tag cfg1x1_tma$tma --streams 1 --steps 200 >> gpurun_out/r02h_kt.jsonl 2>&1
  BFIR_FFT_TMA=$tma python tools/kernel_times.py --tag cfg0_tma$tma --channels 2 --realsize 4 --L 4096 --P 16 --streams 1 --steps 200 >> gpurun_out/r02h_kt.jsonl 2>&1
done
python - <<'PY'
import json
for l in open("gpurun_out/r02h_kt.jsonl"):
    try:
        j = json.loads(l)
        print(j["tag"], {k: round(v, 5) for k, v in j["ms"].items()}, round(j["Msamples_s"]))
    except Exception:
        print(l[:300])
PY
run() {
  echo "== $1"
  env $1 python bench.py --steps 200 --warmup 4 --no-configs --no-latency --no-cpu-baseline > gpurun_out/r02h_tmp.json 2> gpurun_out/r02h_tmp.err || tail -3 gpurun_out/r02h_tmp.err
  python - <<'PY'
import json
j = json.loads(open("gpurun_out/r02h_tmp.json").read().strip().splitlines()[-1])
r = j["roofline"]
print("value %.0f  mac_ms %.4f  serial value %.0f share %s" % (j["value"], r["avg_launch_ms"], r["serial_pass"]["value"], {k: round(v, 4) for k, v in r["step_share"].items()}))
PY
}
run "BFIR_FFT_TMA=0"
run "BFIR_FFT_TMA=1"
python tools/kernel_times.py --tag ncu --steps 8 > gpurun_out/r02h_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rfft_ -s 70 -c 4 -f -o gpurun_out/r02h_fft_f64_13 python tools/kernel_times.py --tag ncu --steps 8 > gpurun_out/r02h_ncu.log 2>&1
tail -2 gpurun_out/r02h_ncu.log
