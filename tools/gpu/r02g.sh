#!/bin/bash
# round 2, GPU call G: does a cluster launch of the transform kernels cost anything (scheduling) in the staged path?
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() {
  echo "== $1"
  env $1 python bench.py --steps 200 --warmup 4 --no-configs --no-latency --no-cpu-baseline > gpurun_out/r02g_tmp.json 2> gpurun_out/r02g_tmp.err || tail -3 gpurun_out/r02g_tmp.err
  python - <<'PY'
import json
j = json.loads(open("gpurun_out/r02g_tmp.json").read().strip().splitlines()[-1])
r = j["roofline"]
print("value %.0f  mac_ms %.4f  serial value %.0f share %s" % (j["value"], r["avg_launch_ms"], r["serial_pass"]["value"], {k: round(v, 4) for k, v in r["step_share"].items()}))
PY
}
run "BFIR_NOP=1"
run "BFIR_FFT_CLUSTER=8"
run "BFIR_FFT_CLUSTER=4"
run "BFIR_FFT_CLUSTER=2"
run "BFIR_NOP=1"
