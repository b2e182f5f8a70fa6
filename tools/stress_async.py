#!/usr/bin/env python
"""Randomised schedule stress of the pipelined entry points against the one-block device path: for random
geometries, every block goes through a randomly chosen call (bfir_run, bfir_run_device, bfir_run_device_pipelined,
bfir_run_device_pair, bfir_run_device_quad (joined / staged), bfir_run_device_oct (joined / staged), bfir_run_async,
bfir_run_async_pair, bfir_run_async_quad) with random waits in between; outputs must match the reference engine (one stream,
bfir_run_device) to rounding. Every fourth geometry is large enough for the eight-block kernel.
python tools/stress_async.py [seconds] [seed]"""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("foo-dsp-bfir_b200")
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
t_end = time.time() + budget
trials = 0
while time.time() < t_end:
    rs = int(rng.choice([4, 8]))
    L = int(rng.choice([64, 256, 1024]))
    P = int(rng.integers(2, 7))
    C = int(rng.integers(1, 4))
    S = int(rng.integers(1, 7))
    G = int(rng.integers(1, 5))
    if trials % 4 == 3:                  # enough work for the one-slice eight-block kernel (one stream group)
        L, C, S, G = 2048, 4, int(rng.integers(20, 28)) * (2 if rs == 4 else 1), 1
    fmt = pkg.FLOAT_LE if rs == 4 else pkg.FLOAT64_LE
    dt, tdt = (np.float32, torch.float32) if rs == 4 else (np.float64, torch.float64)
    ref = pkg.Brutefir(L, P, rs, C, fmt, fmt, 2000, False, n_streams=S, n_groups=1)
    eng = pkg.Brutefir(L, P, rs, C, fmt, fmt, 2000, False, n_streams=S, n_groups=G)
    h = [rng.standard_normal(L * P) * np.exp(-np.arange(L * P) / (L * P / 4)) for _ in range(C * S)]
    assert ref.set_coeff(h, P) == 0 and eng.set_coeff(h, P) == 0
    nblk = int(rng.integers(12, 40))
    n = S * L * C
    blocks = [rng.uniform(-1, 1, n).astype(dt) for _ in range(nblk)]
    d_in = [torch.from_numpy(b).cuda() for b in blocks]
    pin_in = [torch.from_numpy(b).pin_memory() for b in blocks]
    want = [torch.zeros(n, dtype=tdt, device="cuda") for _ in range(nblk)]
    got_d = [torch.zeros(n, dtype=tdt, device="cuda") for _ in range(nblk)]
    got_h = [torch.zeros(n, dtype=tdt).pin_memory() for _ in range(nblk)]
    where = [None] * nblk
    torch.cuda.synchronize()
    for b in range(nblk):
        ref.run_device(d_in[b], want[b])
    assert ref.sync() == 0
    b, tickets, log = 0, [], []
    while b < nblk:
        op = int(rng.integers(0, 10))
        if op in (3, 5) and b + 1 >= nblk:
            op = 0
        if op in (6, 7, 9) and b + 3 >= nblk:
            op = 1
        if op == 8 and b + 7 >= nblk:
            op = 2
        log.append(op)
        if op == 0:
            rc, out = eng.run(blocks[b].view(np.uint8), got_h[b].numpy().view(np.uint8)); assert rc == 0; where[b] = "h"; b += 1
        elif op == 1:
            eng.run_device(d_in[b], got_d[b]); where[b] = "d"; b += 1
        elif op == 2:
            eng.run_device_pipelined(d_in[b], got_d[b]); where[b] = "d"; b += 1
        elif op == 3:
            eng.run_device_pair(d_in[b], d_in[b + 1], got_d[b], got_d[b + 1], pipelined=int(rng.integers(0, 3))); where[b] = where[b + 1] = "d"; b += 2
        elif op in (6, 7):
            eng.run_device_quad(d_in[b:b + 4], got_d[b:b + 4], staged=(op == 7))
            for k in range(4):
                where[b + k] = "d"
            b += 4
        elif op == 8:
            eng.run_device_oct(d_in[b:b + 8], got_d[b:b + 8], staged=bool(rng.integers(0, 2)))
            for k in range(8):
                where[b + k] = "d"
            b += 8
        elif op == 9:
            tickets.append(eng.run_async_quad([p.numpy() for p in pin_in[b:b + 4]], [p.numpy() for p in got_h[b:b + 4]]))
            for k in range(4):
                where[b + k] = "h"
            b += 4
        elif op == 4:
            tickets.append(eng.run_async(pin_in[b].numpy(), got_h[b].numpy())); where[b] = "h"; b += 1
        else:
            tickets.append(eng.run_async_pair(pin_in[b].numpy(), pin_in[b + 1].numpy(), got_h[b].numpy(), got_h[b + 1].numpy())); where[b] = where[b + 1] = "h"; b += 2
        r = rng.random()
        if r < 0.15 and tickets:
            assert eng.wait(tickets[int(rng.integers(0, len(tickets)))]) == 0
        elif r < 0.25:
            assert eng.sync() == 0
        elif r < 0.30:
            eng.join()
    assert eng.sync() == 0
    assert eng.blockcounter() == nblk
    tol = 3e-6 if rs == 4 else 1e-13
    for k in range(nblk):
        a = want[k].cpu().numpy().astype(np.float64)
        g = (got_d[k].cpu().numpy() if where[k] == "d" else got_h[k].numpy()).astype(np.float64)
        err = np.sqrt(np.mean((a - g) ** 2) / max(np.mean(a ** 2), 1e-300))
        assert err < tol, ("mismatch", dict(rs=rs, L=L, P=P, C=C, S=S, G=G), k, err, log)
    eng.close(); ref.close()
    trials += 1
print("stress ok: %d random schedules" % trials)
