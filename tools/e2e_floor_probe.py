#!/usr/bin/env python
"""Where does the end-to-end step lose time against the link? The pinned-host four-block path (bfir_run_async_quad) and
the two-block path at the bench geometry (cfg1 x 16 streams) with P partitions: P = 32 is the bench, P = 2 makes the
partition sum negligible, so what remains is the copy / event structure itself.
    python tools/e2e_floor_probe.py [P ...]"""
import importlib, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("foo-dsp-bfir_b200")
L, C, S, steps = 8192, 8, 16, 400
Ps = [int(x) for x in sys.argv[1:]] or [32, 2]
n_host = 16
host_in = [torch.rand(S * L * C, dtype=torch.float64).pin_memory() for _ in range(n_host)]
host_out = [torch.empty(S * L * C, dtype=torch.float64).pin_memory() for _ in range(n_host)]
ins, outs = [h.numpy() for h in host_in], [h.numpy() for h in host_out]
base = np.random.default_rng(0).standard_normal(L * 32) * np.exp(-6.9 * np.arange(L * 32) / (L * 32))
for P in Ps:
    e = pkg.Brutefir(L, P, 8, C, pkg.FLOAT64_LE, pkg.FLOAT64_LE, 48000, False, n_streams=S, n_groups=1)
    assert e.set_coeff([np.roll(base[:L * P], c) for c in range(S * C)], P) == 0
    for b in range(P + 2):
        e.run(ins[b % n_host], outs[0])
    res = {"P": P}
    for mode, groups, depth in (("quad", 1, 2), ("pair", 4, 3), ("quad", 1, 2), ("quad", 1, 3)):
        e.set_groups(groups)
        per = 4 if mode == "quad" else 2
        times = []
        for rep in range(7):
            tickets = []
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for k in range(steps // per):
                b = per * k
                if mode == "quad":
                    tickets.append(e.run_async_quad([ins[(b + j) % n_host] for j in range(4)], [outs[(b + j) % n_host] for j in range(4)]))
                else:
                    tickets.append(e.run_async_pair(ins[b % n_host], ins[(b + 1) % n_host], outs[b % n_host], outs[(b + 1) % n_host]))
                if k >= depth:
                    assert e.wait(tickets[k - depth]) == 0
            assert e.wait(tickets[-1]) == 0
            torch.cuda.synchronize()
            times.append((time.perf_counter() - t0) / steps * 1e3)
        res["%s_g%d_d%d" % (mode, groups, depth)] = sorted(times[1:])[1]
        res["%s_g%d_d%d_all" % (mode, groups, depth)] = [round(x, 4) for x in times]
    e.close()
    print(json.dumps(res))
