#!/usr/bin/env python
"""Does the host link keep its duplex rate while the partition sums saturate HBM? Copies of 2 MiB (what one stream
group moves per block) in both directions on two streams, alone and with cfg1 x 16 pair steps running on a third."""
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
d_in = [torch.rand(n, dtype=torch.float64, device="cuda") for _ in range(2)]
d_out = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(2)]
for b in range(P + 2):
    e.run_device(d_in[b % 2], d_out[0])
e.sync()
chunk = (int(sys.argv[1]) if len(sys.argv) > 1 else 2) << 20
h_in = torch.empty(chunk, dtype=torch.uint8).pin_memory(); h_out = torch.empty(chunk, dtype=torch.uint8).pin_memory()
g_in = torch.empty(chunk, dtype=torch.uint8, device="cuda"); g_out = torch.empty(chunk, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def copies(iters):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        with torch.cuda.stream(s1):
            g_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(g_out, non_blocking=True)
    s1.synchronize(); s2.synchronize()
    return (time.perf_counter() - t0) / iters


res = {"chunk_MiB": chunk >> 20}
dt = copies(400)
res["alone_GBs_total"] = 2 * chunk / dt / 1e9
for _ in range(300):                       # ~40 ms of partition sums queued on the engine's stream
    e.run_device_pair(d_in[0], d_in[1], d_out[0], d_out[1])
dt = copies(400)
res["under_load_GBs_total"] = 2 * chunk / dt / 1e9
e.sync()
print(json.dumps(res))
