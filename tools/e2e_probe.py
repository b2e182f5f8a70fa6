#!/usr/bin/env python
"""End-to-end step time of cfg1 x 16 streams on pinned host buffers, three ways of driving the engine:

    python tools/e2e_probe.py sync <groups>               one engine, synchronous bfir_run
    python tools/e2e_probe.py async <groups> <depth>      one engine, bfir_run_async with <depth> blocks in flight
    python tools/e2e_probe.py threads <T> <groups>        T engines of 16/T streams, one host thread each, bfir_run

Round-1 findings on the B200 box (16.8 MB over PCIe per step): copies alone 0.319 ms serial / 0.282 ms with
H2D and D2H on different streams (the link gives ~60 GB/s combined, so the two directions barely overlap),
kernels alone 0.226 ms; synchronous: 0.537 ms with 1 group, 0.393 ms with 4, 0.385 ms with 6."""
import importlib, json, os, sys, threading, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("foo-dsp-bfir_b200")
L, P, C, S = 8192, 32, 8, 16
mode = sys.argv[1] if len(sys.argv) > 1 else "sync"
a1 = int(sys.argv[2]) if len(sys.argv) > 2 else 4
a2 = int(sys.argv[3]) if len(sys.argv) > 3 else 2
base = np.random.default_rng(0).standard_normal(L * P) * np.exp(-6.9 * np.arange(L * P) / (L * P))
N_STEPS = 300


def engine(streams, groups):
    e = pkg.Brutefir(L, P, 8, C, pkg.FLOAT64_LE, pkg.FLOAT64_LE, 48000, False, n_streams=streams, n_groups=groups)
    assert e.set_coeff([np.roll(base, c) for c in range(streams * C)], P) == 0
    return e


def pinned(streams, n):
    return [torch.rand(streams * L * C, dtype=torch.float64).pin_memory().numpy() for _ in range(n)]


if mode == "sync":
    e, (a,), (b,) = engine(S, a1), pinned(S, 1), pinned(S, 1)
    for _ in range(P + 5):
        e.run(a, b)
    t0 = time.perf_counter()
    for _ in range(N_STEPS):
        e.run(a, b)
    dt = (time.perf_counter() - t0) / N_STEPS
elif mode == "async":
    depth = a2
    e, ins, outs = engine(S, a1), pinned(S, depth + 1), pinned(S, depth + 1)
    for _ in range(P + 5):
        e.run(ins[0], outs[0])
    assert e.wait(e.run_async(ins[0], outs[0])) == 0      # first use allocates the staging ring: keep it out of the timing
    tickets = []
    t0 = time.perf_counter()
    for k in range(N_STEPS):
        tickets.append(e.run_async(ins[k % (depth + 1)], outs[k % (depth + 1)]))
        if k >= depth:
            assert e.wait(tickets[k - depth]) == 0
    assert e.wait(tickets[-1]) == 0
    dt = (time.perf_counter() - t0) / N_STEPS
else:
    T, groups = a1, a2
    engs = [engine(S // T, groups) for _ in range(T)]
    bufs = [(pinned(S // T, 1)[0], pinned(S // T, 1)[0]) for _ in range(T)]
    for e, (a, b) in zip(engs, bufs):
        for _ in range(P + 5):
            e.run(a, b)
    start = threading.Barrier(T + 1)

    def work(i):
        e, (a, b) = engs[i], bufs[i]
        start.wait()
        for _ in range(N_STEPS):
            e.run(a, b)
    th = [threading.Thread(target=work, args=(i,)) for i in range(T)]
    for t in th:
        t.start()
    start.wait()
    t0 = time.perf_counter()
    for t in th:
        t.join()
    dt = (time.perf_counter() - t0) / N_STEPS
print(json.dumps({"mode": mode, "args": [a1, a2], "ms_per_step": dt * 1e3, "msamples_per_s": S * C * L / dt / 1e6}))
