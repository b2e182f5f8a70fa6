#!/bin/bash
# round 2, GPU call F: A/B of co-residency variants for the staged four-block path (cfg1 x 16)
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "quad or stage" > gpurun_out/r02f_pytest.log 2>&1; tail -2 gpurun_out/r02f_pytest.log
run() {
  echo "== $1"
  env $1 python bench.py --steps 200 --warmup 4 --no-configs --no-latency --no-cpu-baseline > gpurun_out/r02f_tmp.json 2> gpurun_out/r02f_tmp.err || tail -3 gpurun_out/r02f_tmp.err
  python - <<'PY'
import json
j = json.loads(open("gpurun_out/r02f_tmp.json").read().strip().splitlines()[-1])
r = j["roofline"]
print("value %.0f  mac_ms %.4f  serial value %.0f share %s" % (j["value"], r["avg_launch_ms"], r["serial_pass"]["value"], {k: round(v, 4) for k, v in r["step_share"].items()}))
PY
}
run "BFIR_NOP=1"
run "BFIR_QUAD_THREADS=128"
run "BFIR_FFT_R0=2"
run "BFIR_FFT_R0=2 BFIR_QUAD_THREADS=128"
run "BFIR_FFT_R0=2 BFIR_FFT_E=8"
run "BFIR_FFT_R0=2 BFIR_FFT_E=8 BFIR_QUAD_THREADS=128"
run "BFIR_FFT_E=8"
