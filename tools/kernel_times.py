#!/usr/bin/env python
"""Per-kernel CUDA-event times of the engine's block step for an arbitrary configuration (tuning aid).

    python tools/kernel_times.py --channels 8 --realsize 8 --L 8192 --P 32 --streams 16 [--steps 200]
Prints one JSON line: ms per step for {fwd, mac, inv}, MAC GB/s against B_mac, Msamples/s.
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--channels", type=int, default=8)
    ap.add_argument("--realsize", type=int, default=8)
    ap.add_argument("--L", type=int, default=8192)
    ap.add_argument("--P", type=int, default=32)
    ap.add_argument("--streams", type=int, default=16)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--tag", default="")
    ap.add_argument("--quads", action="store_true", help="four blocks per call (bfir_run_device_quad, single precision)")
    ap.add_argument("--octs", action="store_true", help="eight blocks per call (bfir_run_device_oct, joined): only the partition-sum time is meaningful")
    ap.add_argument("--pairs", action="store_true", help="two blocks per call (bfir_run_device_pair); times are per PAIR then")
    a = ap.parse_args()
    import torch
    pkg = importlib.import_module("foo-dsp-bfir_b200")
    fmt = pkg.FLOAT_LE if a.realsize == 4 else pkg.FLOAT64_LE
    dt = torch.float32 if a.realsize == 4 else torch.float64
    eng = pkg.Brutefir(a.L, a.P, a.realsize, a.channels, fmt, fmt, 48000, False, n_streams=a.streams, n_groups=1)
    Ct = a.streams * a.channels
    rng = np.random.default_rng(0)
    taps = a.L * a.P
    env = np.exp(-6.9 * np.arange(taps) / taps)
    base = rng.standard_normal(taps) * env
    base /= np.sqrt(np.sum(base * base))
    # distinct filters per channel without generating Ct x taps random numbers
    hs = [np.roll(base, c % 97) * (1.0 + 0.001 * (c % 13)) for c in range(Ct)]
    assert eng.set_coeff(hs, a.P) == 0
    d_in = [torch.rand(a.streams * a.L * a.channels, dtype=dt, device="cuda") * 2 - 1 for _ in range(4)]
    d_out = torch.empty(a.streams * a.L * a.channels, dtype=dt, device="cuda")
    torch.cuda.synchronize()
    for b in range(a.P + 5):
        eng.run_device(d_in[b % 4], d_out)
    assert eng.sync() == 0
    if a.octs:
        outs = [d_out] + [torch.empty_like(d_out) for _ in range(7)]
        eng.run_device_oct(d_in + d_in, outs)
        assert eng.sync() == 0
        eng.get_mac_profile(8)
        eng.set_profiling(a.steps // 8)
        for b in range(0, a.steps, 8):
            eng.run_device_oct(d_in + d_in, outs)
        assert eng.sync() == 0
        ms8, n8 = eng.get_mac_profile(8)
        Ct_, N_ = a.streams * a.channels, 2 * a.L
        need = (2 * a.P + 15) * N_ * a.realsize * Ct_
        print(json.dumps({"tag": a.tag, "oct_w": os.environ.get("BFIR_OCT_W"), "mac_ms_per_launch": ms8 / max(n8, 1), "mac_ms_per_block": ms8 / max(n8, 1) / 8,
                          "launches": n8, "frac_of_6545_on_bytes_needed": need / (ms8 / max(n8, 1) * 1e-3) / 1e9 / 6545.3}))
        return
    if a.quads:
        outs = [d_out] + [torch.empty_like(d_out) for _ in range(3)]
        eng.run_device_quad(d_in, outs)
        assert eng.sync() == 0
        eng.set_profiling(a.steps // 4)
        for b in range(0, a.steps, 4):
            eng.run_device_quad(d_in, outs)
    elif a.pairs:
        d_out2 = torch.empty_like(d_out)
        eng.run_device_pair(d_in[0], d_in[1], d_out, d_out2)
        assert eng.sync() == 0
        eng.set_profiling(a.steps // 2)
        for b in range(0, a.steps, 2):
            eng.run_device_pair(d_in[b % 4], d_in[(b + 1) % 4], d_out, d_out2)
    else:
        eng.set_profiling(a.steps)
        for b in range(a.steps):
            eng.run_device(d_in[b % 4], d_out)
    assert eng.sync() == 0
    prof, n = eng.get_profile()
    per = 4 if a.quads else (2 if a.pairs else 1)   # blocks per profiled entry
    ms = {k: v / n / per for k, v in prof.items()}   # per block
    step = sum(ms.values())
    b_mac = (2 * a.P + 1) * 2 * a.L * a.realsize * Ct
    print(json.dumps({"tag": a.tag, "split_env": os.environ.get("BFIR_MAC_SPLIT"), "cfg": vars(a), "ms": ms,
                      "step_ms": step, "mac_GBs": b_mac / (ms["mac_ms"] * 1e-3) / 1e9,
                      "mac_frac_of_6545": b_mac / (ms["mac_ms"] * 1e-3) / 1e9 / 6545.3,
                      "Msamples_s": Ct * a.L / (step * 1e-3) / 1e6}))


if __name__ == "__main__":
    main()
