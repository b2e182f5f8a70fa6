#!/usr/bin/env python
"""A few two-block steps (or, with --quads, four-block steps) of cfg1 x 16 streams on one stream, for ncu
(profiles/README.md):
ncu --set full --clock-control none -k regex:partition_mac_pair -s 2 -c 2 python tools/pair_ncu_target.py
ncu --set full --clock-control none -k regex:partition_mac_multi -s 2 -c 2 python tools/pair_ncu_target.py --quads
ncu --set full --clock-control none -k regex:partition_mac_oct -s 2 -c 2 python tools/pair_ncu_target.py --octs"""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("foo-dsp-bfir_b200")
L, P, C, S = 8192, 32, 8, 16
e = pkg.Brutefir(L, P, 8, C, pkg.FLOAT64_LE, pkg.FLOAT64_LE, 48000, False, n_streams=S, n_groups=1)
base = np.random.default_rng(0).standard_normal(L * P) * np.exp(-6.9 * np.arange(L * P) / (L * P))
assert e.set_coeff([np.roll(base, c) for c in range(S * C)], P) == 0
n = S * L * C
d_in = [torch.rand(n, dtype=torch.float64, device="cuda") for _ in range(4)]
d_out = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(4)]
for b in range(P + 2):
    e.run_device(d_in[b % 4], d_out[0])
if "--octs" in sys.argv:
    d_out8 = d_out + [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(4)]
    for b in range(0, 48, 8):
        e.run_device_oct(d_in + d_in, d_out8)
elif "--quads" in sys.argv:
    for b in range(0, 24, 4):
        e.run_device_quad(d_in, d_out)
else:
    for b in range(0, 12, 2):
        e.run_device_pair(d_in[b % 4], d_in[(b + 1) % 4], d_out[0], d_out[1])
assert e.sync() == 0
print("ok")
