#!/usr/bin/env python
"""Block pairs (one partition-sum launch per two blocks) against single blocks, cfg1 x 16 streams:
device-resident pipelined and end to end on pinned host buffers.  python tools/pair_probe.py [groups] [depth_pairs]"""
import importlib, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("foo-dsp-bfir_b200")
L, P, C, S = 8192, 32, 8, 16
groups = int(sys.argv[1]) if len(sys.argv) > 1 else 4
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 2
fmt = int(sys.argv[3]) if len(sys.argv) > 3 else pkg.FLOAT64_LE
tdt = torch.float64 if fmt == pkg.FLOAT64_LE else torch.float32
base = np.random.default_rng(0).standard_normal(L * P) * np.exp(-6.9 * np.arange(L * P) / (L * P))
e = pkg.Brutefir(L, P, 8, C, fmt, fmt, 48000, False, n_streams=S, n_groups=groups)
assert e.set_coeff([np.roll(base, c) for c in range(S * C)], P) == 0
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
e.set_stream(st.cuda_stream)
n = S * L * C
d_in = [torch.rand(n, dtype=tdt, device="cuda") for _ in range(4)]
d_out = [torch.empty(n, dtype=tdt, device="cuda") for _ in range(2)]
for b in range(P + 4):
    e.run_device(d_in[b % 4], d_out[0])
assert e.sync() == 0
K = 200
res = {"groups": e.get_groups(), "depth_pairs": depth}
for name in ("single", "pair"):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(st)
    for b in range(0, K, 2):
        if name == "single":
            e.run_device_pipelined(d_in[b % 4], d_out[0]); e.run_device_pipelined(d_in[(b + 1) % 4], d_out[1])
        else:
            e.run_device_pair(d_in[b % 4], d_in[(b + 1) % 4], d_out[0], d_out[1], pipelined="staged")
    e.join()
    ev1.record(st)
    assert e.sync() == 0
    res["device_%s_ms_per_block" % name] = ev0.elapsed_time(ev1) / K
R = 2 * (depth + 1)
ins = [torch.rand(n, dtype=tdt).pin_memory().numpy() for _ in range(R)]
outs = [torch.empty(n, dtype=tdt).pin_memory().numpy() for _ in range(R)]
assert e.wait(e.run_async(ins[0], outs[0])) == 0          # first use allocates the staging ring: keep it out of the timing
assert e.wait(e.run_async(ins[1], outs[1])) == 0
for name in ("single", "pair"):
    tickets = []
    issue = 0.0
    t0 = time.perf_counter()
    for k in range(K // 2):
        i0, i1 = (2 * k) % R, (2 * k + 1) % R
        ti = time.perf_counter()
        if name == "single":
            e.run_async(ins[i0], outs[i0]); tickets.append(e.run_async(ins[i1], outs[i1]))
        else:
            tickets.append(e.run_async_pair(ins[i0], ins[i1], outs[i0], outs[i1]))
        issue += time.perf_counter() - ti
        if k >= depth:
            assert e.wait(tickets[k - depth]) == 0
    assert e.wait(tickets[-1]) == 0
    dt = (time.perf_counter() - t0) / K
    res["e2e_%s_ms_per_block" % name] = dt * 1e3
    res["e2e_%s_issue_ms_per_block" % name] = issue / K * 1e3
    res["e2e_%s_msamples" % name] = S * C * L / dt / 1e6
print(json.dumps(res))
