#!/usr/bin/env python
"""End-to-end step time of bfir_run on pinned host buffers for cfg1 x 16 streams as a function of the number
of stream groups:  python tools/e2e_probe.py <groups>.
Round-1 findings on the B200 box (16.8 MB over PCIe per step): copies alone 0.319 ms serial / 0.282 ms with
H2D and D2H on different streams (the link gives ~60 GB/s combined, so the two directions barely overlap),
kernels alone 0.226 ms, everything: 0.537 ms with 1 group, 0.393 ms with 4, 0.385 ms with 6 -- the step is
PCIe-bound; issue order (all fronts then all backs vs group by group) makes no difference."""
import importlib, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("foo-dsp-bfir_b200")
L, P, C, S = 8192, 32, 8, 16
groups = int(sys.argv[1]) if len(sys.argv) > 1 else 4
eng = pkg.Brutefir(L, P, 8, C, pkg.FLOAT64_LE, pkg.FLOAT64_LE, 48000, False, n_streams=S, n_groups=groups)
base = np.random.default_rng(0).standard_normal(L * P) * np.exp(-6.9 * np.arange(L * P) / (L * P))
assert eng.set_coeff([np.roll(base, c) for c in range(S * C)], P) == 0
hin = torch.rand(S * L * C, dtype=torch.float64).pin_memory()
hout = torch.empty(S * L * C, dtype=torch.float64).pin_memory()
a, b = hin.numpy(), hout.numpy()
for _ in range(P + 5):
    eng.run(a, b)
t0 = time.perf_counter()
n = 300
for _ in range(n):
    eng.run(a, b)
dt = (time.perf_counter() - t0) / n
print(json.dumps({"groups": groups, "ms_per_step": dt * 1e3}))
