#!/usr/bin/env python
"""Host-visible per-block latency of the single-stream configurations (BASELINE configs[0], [1], [2]):
wall time of the call sequence a real-time host makes for one block, pinned host buffers, p50 / p99.

    cfg0  stereo 44.1 kHz float, 65536 taps, L 4096, P 16:            bfir_run
    cfg1  7.1 48 kHz double, 262144 taps, L 8192, P 32:               bfir_run
    cfg2  stereo 96 kHz float, 131072 taps from the equalizer, swapped in with a crossfade EVERY block:
          bfir_eq_render_device (262144-point four-step inverse FFT) + bfir_set_coeff_device(crossfade)
          (32 partition FFTs) + bfir_run (two partition sums, two inverse FFTs, ramp)
"""
import importlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

pkg = importlib.import_module("foo-dsp-bfir_b200")
BANDS = [20, 25, 31.5, 40, 50, 63, 80, 100, 125, 160, 200, 250, 315, 400, 500, 630, 800, 1000, 1250, 1600, 2000, 2500,
         3150, 4000, 5000, 6300, 8000, 10000, 12500, 16000, 20000]


def stats(lat, period_ms):
    lat = np.sort(np.array(lat)) * 1e3
    return {"p50_ms": float(lat[len(lat) // 2]), "p99_ms": float(lat[int(len(lat) * 0.99)]), "max_ms": float(lat[-1]),
            "calls": len(lat), "block_period_ms": period_ms, "real_time_margin_x": period_ms / float(lat[int(len(lat) * 0.99)])}


def filt(ch, taps):
    g = np.random.default_rng(1000 + ch).standard_normal(taps) * np.exp(-6.9 * np.arange(taps) / taps)
    return g / np.sqrt(np.sum(g * g))


def plain(name, L, P, rs, C, rate, calls=3000):
    fmt = pkg.FLOAT_LE if rs == 4 else pkg.FLOAT64_LE
    dt = torch.float32 if rs == 4 else torch.float64
    e = pkg.Brutefir(L, P, rs, C, fmt, fmt, rate, False)
    assert e.set_coeff([filt(c, L * P) for c in range(C)], P) == 0
    x = [(torch.rand(L * C, dtype=dt) * 2 - 1).pin_memory() for _ in range(4)]
    y = torch.empty(L * C, dtype=dt).pin_memory()
    xs, yn = [t.numpy() for t in x], y.numpy()
    for b in range(P + 50):
        e.run(xs[b % 4], yn)
    out = {}
    # "paced": a real-time host calls once per block period; 300 us of idle time between calls (a small fraction
    # of any block period here) lets the engine's look-ahead partition sum for the next block finish, as it
    # would in use. "back_to_back": calls with no gap, where that work is still on the critical path.
    for mode, gap in (("paced", 300e-6), ("back_to_back", 0.0)):
        lat = []
        for b in range(calls):
            if gap:
                t1 = time.perf_counter() + gap
                while time.perf_counter() < t1:
                    pass
            t0 = time.perf_counter()
            e.run(xs[b % 4], yn)
            lat.append(time.perf_counter() - t0)
        out[mode] = stats(lat, 1e3 * L / rate)
    return dict(out["paced"], back_to_back={k: out["back_to_back"][k] for k in ("p50_ms", "p99_ms", "max_ms")}, pacing_gap_ms=0.3, config=name)


def cfg2(calls=1500):
    L, EQB, C, rs, rate = 4096, 64, 2, 4, 96000
    taps = L * EQB
    P = (taps // 2) // L
    eq = pkg.Equalizer(L, EQB, rs, rate)
    e = pkg.Brutefir(L, P, rs, C, pkg.FLOAT_LE, pkg.FLOAT_LE, rate, False)
    rng = np.random.default_rng(7)
    zero = [0.0] * 31
    assert e.set_coeff_device(eq.generate_device(BANDS, list(rng.integers(-120, 121, 31) / 10.0), zero), 0, C, taps // 2, P) == 0
    gains = [list(rng.integers(-120, 121, 31) / 10.0) for _ in range(16)]
    x = [(torch.rand(L * C) * 2 - 1).pin_memory() for _ in range(4)]
    y = torch.empty(L * C).pin_memory()
    xs, yn = [t.numpy() for t in x], y.numpy()
    lat = []
    for b in range(P + 50 + calls):
        t0 = time.perf_counter()
        d = eq.generate_device(BANDS, gains[b % 16], zero)
        assert e.set_coeff_device(d, 0, C, taps // 2, P, crossfade=True) == 0
        e.run(xs[b % 4], yn)
        if b >= P + 50:
            lat.append(time.perf_counter() - t0)
    return dict(stats(lat, 1e3 * L / rate), config="cfg2: EQ render + crossfade swap + run, every block")


if __name__ == "__main__":
    for r in (plain("cfg0: stereo float 65536 taps", 4096, 16, 4, 2, 44100),
              plain("cfg1: 7.1 double 262144 taps", 8192, 32, 8, 8, 48000),
              plain("product: stereo, REALSIZE 8, FILTER_LEN 1024, 64 blocks", 1024, 64, 8, 2, 44100),
              cfg2()):
        print(json.dumps(r))
