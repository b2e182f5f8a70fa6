#!/usr/bin/env python
"""Does the engine's kernel load slow the host link down? The bench's link-only copies (8 MiB H2D + 8 MiB D2H per step on two
streams, no dependencies) timed alone and while the device-resident eight-block stage pipeline of cfg1 x 16 runs beside them.
    python tools/link_under_kernels.py"""
import importlib, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("foo-dsp-bfir_b200")
L, P, C, S = 8192, 32, 8, 16
e = pkg.Brutefir(L, P, 8, C, pkg.FLOAT64_LE, pkg.FLOAT64_LE, 48000, False, n_streams=S, n_groups=1)
base = np.random.default_rng(0).standard_normal(L * P) * np.exp(-6.9 * np.arange(L * P) / (L * P))
assert e.set_coeff([np.roll(base, c) for c in range(S * C)], P) == 0
n = S * L * C
d_in = [torch.rand(n, dtype=torch.float64, device="cuda") for _ in range(8)]
d_out = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(8)]
for b in range(P + 2):
    e.run_device(d_in[b % 8], d_out[0])
e.run_device_oct(d_in, d_out, staged=True); e.join(); assert e.sync() == 0
NH = 12
h_in = [torch.rand(n, dtype=torch.float64).pin_memory() for _ in range(NH)]
h_out = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(NH)]
c_in = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(2)]
c_out = [torch.rand(n, dtype=torch.float64, device="cuda") for _ in range(2)]
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()


def copies(steps, kernels):
    torch.cuda.synchronize()
    if kernels:                                   # ~0.5 ms of GPU work per call, queued ahead of the copies
        for k in range(kernels):
            e.run_device_oct(d_in, d_out, staged=True)
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record(s_in); b0.record(s_out)
    for b in range(steps):
        with torch.cuda.stream(s_in):
            c_in[b & 1].copy_(h_in[b % NH], non_blocking=True)
        with torch.cuda.stream(s_out):
            h_out[b % NH].copy_(c_out[b & 1], non_blocking=True)
    a1.record(s_in); b1.record(s_out)
    torch.cuda.synchronize()
    if kernels:
        e.join(); assert e.sync() == 0
    return a0.elapsed_time(a1) / steps, b0.elapsed_time(b1) / steps


copies(50, 0)
alone = [copies(200, 0) for _ in range(3)]
loaded = [copies(200, 90) for _ in range(3)]     # 90 calls x ~0.5 ms cover the 200 x 0.18 ms of copies
print(json.dumps({"alone_ms_per_step_h2d_d2h": alone, "beside_the_stage_pipeline": loaded}))
