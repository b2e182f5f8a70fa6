"""GPU parity of the crossbar engine (BASELINE configs[4]: "32x32 mixnscale crossbar"): the input
and output gain matrices are convolver_mixnscale with n_bufs > 1 (reference
brutefir/fftw_convolver.cpp:215-229) around the per-filter partition sums. The oracle side composes
the reference's own entry points in the order of brutefir::run."""
import numpy as np
import pytest

from conftest import white_noise, decay_filter, rel_rms

pytestmark = pytest.mark.gpu


def oracle_xbar_run(oracle, L, P, rs, n_in, n_f, n_out, h, gin, gout, x):
    """x: [blocks*L, n_in] -> y: [blocks*L, n_out] using fftw_convolver entry points only"""
    cv = oracle.Convolver(L, rs)
    dt = cv.dtype
    H = [cv.preprocess_coeff(np.asarray(h[f], dtype=dt), P) for f in range(n_f)]
    fdl = np.zeros((n_f, P, 2 * L), dtype=dt)
    prev = np.zeros((n_in, L), dtype=dt)
    nb = x.shape[0] // L
    y = np.zeros((nb * L, n_out))
    for t in range(nb):
        hcs = []
        for i in range(n_in):
            cur = x[t * L:(t + 1) * L, i].astype(dt)
            hcs.append(cv.time2freq(np.concatenate([prev[i], cur])))
            prev[i] = cur
        accs = []
        for f in range(n_f):
            fdl[f, t % P] = cv.mixnscale(hcs, list(gin[f]), 1)                  # MIXMODE_INPUT, n_bufs = n_in
            acc = cv.convolve(fdl[f, t % P].copy(), H[f][0].copy())
            for i in range(1, min(P, t + 1)):
                cv.convolve_add(fdl[f, (t - i) % P].copy(), H[f][i].copy(), acc)
            accs.append(acc)
        for o in range(n_out):
            hc = cv.mixnscale(accs, list(gout[o]), 3)                           # MIXMODE_OUTPUT, n_bufs = n_f
            y[t * L:(t + 1) * L, o] = cv.freq2time(hc)[:L]
    return y


@pytest.mark.parametrize("rs,L,P,n_in,n_f,n_out", [(4, 256, 3, 2, 3, 2), (8, 128, 4, 5, 4, 3), (4, 1024, 2, 32, 32, 32), (8, 64, 2, 1, 6, 2)])
def test_crossbar_matches_mixnscale_composition(pkg, oracle, rs, L, P, n_in, n_f, n_out):
    fmt = pkg.FLOAT_LE if rs == 4 else pkg.FLOAT64_LE
    dt = np.float32 if rs == 4 else np.float64
    rng = np.random.default_rng(n_in * 100 + n_f)
    gin = rng.standard_normal((n_f, n_in)) / np.sqrt(n_in)
    gout = rng.standard_normal((n_out, n_f)) / np.sqrt(n_f)
    h = [decay_filter(f, L * P) for f in range(n_f)]
    g = pkg.Brutefir(L, P, rs, n_f, fmt, fmt, 48000, False, xbar_inputs=n_in, xbar_outputs=n_out)
    assert g.set_coeff(h, P) == 0
    x = white_noise(21, (2 * P + 1) * L, n_in).astype(dt)
    with pytest.raises(pkg.BfirError):
        g.run(np.ascontiguousarray(x[:L]).view(np.uint8).ravel())          # crossbar gains not set yet
    g.set_crossbar(gin, gout)
    ys = []
    for b in range(2 * P + 1):
        rc, out = g.run(np.ascontiguousarray(x[b * L:(b + 1) * L]).view(np.uint8).ravel())
        assert rc == 0
        ys.append(out.view(dt).reshape(L, n_out).copy())
    y = np.concatenate(ys)
    ref = oracle_xbar_run(oracle, L, P, rs, n_in, n_f, n_out, h, gin if rs == 8 else gin.astype(np.float32), gout if rs == 8 else gout.astype(np.float32), x)
    for o in range(n_out):
        assert rel_rms(y[:, o], ref[:, o]) < (1e-5 if rs == 4 else 1e-12)


def test_identity_crossbar_equals_diagonal_engine(pkg, monkeypatch):
    L, P, C = 512, 3, 4
    h = [decay_filter(c, L * P) for c in range(C)]
    monkeypatch.setenv("BFIR_LOOKAHEAD", "0")    # one-kernel partition sum on both engines (the crossbar engine has no look-ahead)
    a = pkg.Brutefir(L, P, 4, C, pkg.FLOAT_LE, pkg.FLOAT_LE, 48000, False)
    b = pkg.Brutefir(L, P, 4, C, pkg.FLOAT_LE, pkg.FLOAT_LE, 48000, False, xbar_inputs=C, xbar_outputs=C)
    a.set_coeff(h, P); b.set_coeff(h, P)
    b.set_crossbar(np.eye(C), np.eye(C))
    x = white_noise(3, 7 * L, C).astype(np.float32)
    for blk in range(7):
        raw = np.ascontiguousarray(x[blk * L:(blk + 1) * L]).view(np.uint8).ravel()
        ya, yb = a.run(raw)[1].view(np.float32), b.run(raw)[1].view(np.float32)
        assert np.array_equal(ya, yb)          # gain 1.0 / 0.0 FMAs are exact


def test_cfg4_shape_impulse_filters_exact(pkg):
    """BASELINE configs[4] geometry on one GPU at reduced depth: 32 inputs x 32 filters x 32 outputs,
    L = 32768 (FFT 65536, two-CTA transform), P = 4. Filters are unit impulses at known delays, so the
    output is a delayed gain-matrix mix of the inputs -- checked without any oracle."""
    L, P, n = 32768, 4, 32
    rng = np.random.default_rng(4)
    delays = rng.integers(0, L * P - 1, n)
    h = []
    for f in range(n):
        v = np.zeros(L * P, dtype=np.float32)
        v[delays[f]] = 1.0
        h.append(v)
    gin = rng.standard_normal((n, n)) / np.sqrt(n)
    gout = rng.standard_normal((n, n)) / np.sqrt(n)
    g = pkg.Brutefir(L, P, 4, n, pkg.FLOAT_LE, pkg.FLOAT_LE, 48000, False, xbar_inputs=n, xbar_outputs=n)
    assert g.set_coeff(h, P) == 0
    g.set_crossbar(gin, gout)
    nb = P + 2
    x = white_noise(8, nb * L, n).astype(np.float32)
    y = np.concatenate([g.run(np.ascontiguousarray(x[b * L:(b + 1) * L]).view(np.uint8).ravel())[1].view(np.float32).reshape(L, n).copy()
                        for b in range(nb)])
    fin = x.astype(np.float64) @ gin.astype(np.float32).astype(np.float64).T          # filter inputs [t, f]
    fout = np.zeros_like(fin)
    for f in range(n):
        d = delays[f]
        fout[d:, f] = fin[: nb * L - d, f]
    ref = fout @ gout.astype(np.float32).astype(np.float64).T
    assert rel_rms(y, ref) < 1e-5


def test_cfg4_full_depth_one_gpu(pkg):
    """BASELINE configs[4] at FULL size on one GPU: 16 Mi taps (L 32768 x P 512) per filter, 32x32x32
    crossbar: 4 GiB of coefficient spectra + 4 GiB delay line. Filters are a few scaled impulses spread
    over the observable part of the 16 Mi taps, so the exact output is a sum of delayed crossbar mixes."""
    import torch
    L, P, n = 32768, 512, 32
    rng = np.random.default_rng(12)
    taps = L * P
    nb = 24                                            # 786432 frames: delays below that are observable
    delays = rng.integers(0, nb * L - 1, (n, 3))
    amps = rng.uniform(0.3, 1.0, (n, 3)).astype(np.float32)
    coeffs = []
    for f in range(n):
        v = np.zeros(taps, dtype=np.float32)
        for d, a in zip(delays[f], amps[f]):
            v[d] += a
        coeffs.append(v)
    gin = rng.standard_normal((n, n)) / np.sqrt(n)
    gout = rng.standard_normal((n, n)) / np.sqrt(n)
    g = pkg.Brutefir(L, P, 4, n, pkg.FLOAT_LE, pkg.FLOAT_LE, 48000, False, xbar_inputs=n, xbar_outputs=n)
    assert g.set_coeff(coeffs, P) == 0
    del coeffs
    g.set_crossbar(gin, gout)
    gen = torch.Generator(device="cuda").manual_seed(9)
    x = [torch.rand(L, n, dtype=torch.float32, device="cuda", generator=gen) * 2 - 1 for _ in range(nb)]
    y = [torch.empty(L, n, dtype=torch.float32, device="cuda") for _ in range(nb)]
    torch.cuda.synchronize()                           # the engine runs on its own stream
    for b in range(nb):
        g.run_device(x[b], y[b])
    assert g.sync() == 0
    X = torch.cat(x, dim=0).double()                                   # [T, n]
    Gi = torch.from_numpy(gin.astype(np.float32)).double().cuda()
    Go = torch.from_numpy(gout.astype(np.float32)).double().cuda()
    fin = X @ Gi.T                                                     # filter inputs [T, f]
    fout = torch.zeros_like(fin)
    for f in range(n):
        for d, a in zip(delays[f], amps[f]):
            d = int(d)
            fout[d:, f] += fin[: nb * L - d, f] * float(a)
    ref = fout @ Go.T
    Y = torch.cat(y, dim=0).double()
    err = torch.sqrt(torch.mean((Y - ref) ** 2) / torch.mean(ref ** 2)).item()
    assert err < 1e-5, err


def test_cfg4_steady_state_all_512_partitions(pkg):
    """BASELINE configs[4] at FULL size in STEADY STATE: 518 blocks, so that from block 511 on all 512 partitions of
    every filter are live (procblocks == P; delay-line slots taken modulo P + 3 over more than a full turn). Filter f
    carries 16 scaled impulses, one in each of the partitions f, f + 32, ..., f + 480 -- over the 32 filters every
    partition index is hit -- so the exact output is a sum of delayed crossbar mixes (float64 on the GPU). Checked on
    the last six blocks: two one-block steps, then two two-block (pair-kernel) steps."""
    import torch
    L, P, n = 32768, 512, 32
    rng = np.random.default_rng(21)
    taps = L * P
    nb = P + 6
    delays = np.array([[(f + 32 * j) * L + int(rng.integers(0, L)) for j in range(16)] for f in range(n)])
    amps = rng.uniform(0.3, 1.0, (n, 16)).astype(np.float32)
    coeffs = []
    for f in range(n):
        v = np.zeros(taps, dtype=np.float32)
        v[delays[f]] = amps[f]
        coeffs.append(v)
    gin = rng.standard_normal((n, n)) / np.sqrt(n)
    gout = rng.standard_normal((n, n)) / np.sqrt(n)
    g = pkg.Brutefir(L, P, 4, n, pkg.FLOAT_LE, pkg.FLOAT_LE, 48000, False, xbar_inputs=n, xbar_outputs=n)
    assert g.set_coeff(coeffs, P) == 0
    del coeffs
    g.set_crossbar(gin, gout)
    gen = torch.Generator(device="cuda").manual_seed(10)
    X = torch.rand(nb * L, n, dtype=torch.float32, device="cuda", generator=gen) * 2 - 1
    Y = torch.empty(6 * L, n, dtype=torch.float32, device="cuda")
    scratch = torch.empty(L, n, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()                           # the engine runs on its own stream: X must be complete first
    first = nb - 6
    for b in range(first + 2):
        g.run_device(X[b * L:(b + 1) * L], Y[(b - first) * L:(b - first + 1) * L] if b >= first else scratch)
    for b in range(first + 2, nb, 2):
        k = b - first
        g.run_device_pair(X[b * L:(b + 1) * L], X[(b + 1) * L:(b + 2) * L], Y[k * L:(k + 1) * L], Y[(k + 1) * L:(k + 2) * L])
    assert g.sync() == 0
    assert g.blockcounter() == nb
    Gi = torch.from_numpy(gin.astype(np.float32)).double().cuda()
    Go = torch.from_numpy(gout.astype(np.float32)).double().cuda()
    m0, m1 = first * L, nb * L
    fout = torch.zeros(m1 - m0, n, dtype=torch.float64, device="cuda")
    for f in range(n):
        for d, a in zip(delays[f], amps[f]):
            d = int(d)
            lo = max(m0 - d, 0)                                     # samples before the start of the run are zero
            seg = X[lo:m1 - d].double() @ Gi[f]                      # filter input f over the shifted window
            fout[(lo + d) - m0:, f] += seg * float(a)
    ref = fout @ Go.T
    for k in range(6):
        a, b = Y[k * L:(k + 1) * L].double(), ref[k * L:(k + 1) * L]
        err = torch.sqrt(torch.mean((a - b) ** 2) / torch.mean(b ** 2)).item()
        print("cfg4 steady state, block %d (%s): rel rms vs float64 truth %.3e" % (first + k, "single" if k < 2 else "pair", err))
        assert err < 1e-5, (k, err)


@pytest.mark.parametrize("rs,L,P,n_in,n_f,n_out,S", [(4, 256, 3, 2, 3, 2, 1), (8, 128, 4, 5, 4, 3, 2), (4, 1024, 4, 32, 32, 32, 1)])
def test_crossbar_block_pairs_equal_single_blocks(pkg, rs, L, P, n_in, n_f, n_out, S):
    """two blocks per partition-sum launch with a crossbar around the filters (the cfg4 shape): same output as block
    by block up to the summation order, through the device and the pinned-host entry points"""
    import torch
    fmt = pkg.FLOAT_LE if rs == 4 else pkg.FLOAT64_LE
    dt, tdt = (np.float32, torch.float32) if rs == 4 else (np.float64, torch.float64)
    rng = np.random.default_rng(5)
    gin = rng.standard_normal((n_f, n_in)) / np.sqrt(n_in)
    gout = rng.standard_normal((n_out, n_f)) / np.sqrt(n_f)
    h = [decay_filter(f % 5, L * P) * (1 + 0.1 * f) for f in range(n_f * S)]
    engines = [pkg.Brutefir(L, P, rs, n_f, fmt, fmt, 48000, False, n_streams=S, xbar_inputs=n_in, xbar_outputs=n_out) for _ in range(3)]
    for e in engines:
        assert e.set_coeff(h, P) == 0
        e.set_crossbar(gin, gout)
    single, pair_d, pair_h = engines
    nblk = 2 * P + 6
    blocks = [rng.uniform(-1, 1, S * L * n_in).astype(dt) for _ in range(nblk)]
    d_in = [torch.from_numpy(b).cuda() for b in blocks]
    pin_in = [torch.from_numpy(b).pin_memory() for b in blocks]
    n_o = S * L * n_out
    out_s = [torch.zeros(n_o, dtype=tdt, device="cuda") for _ in range(nblk)]
    out_p = [torch.zeros(n_o, dtype=tdt, device="cuda") for _ in range(nblk)]
    out_h = [torch.zeros(n_o, dtype=tdt).pin_memory() for _ in range(nblk)]
    torch.cuda.synchronize()
    for b in range(nblk):
        single.run_device(d_in[b], out_s[b])
    t = None
    for b in range(0, nblk, 2):
        pair_d.run_device_pair(d_in[b], d_in[b + 1], out_p[b], out_p[b + 1])
        t = pair_h.run_async_pair(pin_in[b].numpy(), pin_in[b + 1].numpy(), out_h[b].numpy(), out_h[b + 1].numpy())
    assert single.sync() == 0 and pair_d.sync() == 0 and pair_h.wait(t) == 0
    tol = 2e-6 if rs == 4 else 1e-13
    for b in range(nblk):
        a = out_s[b].cpu().numpy()
        assert rel_rms(out_p[b].cpu().numpy(), a) < tol, b
        assert rel_rms(out_h[b].numpy(), a) < tol, b
    # four blocks per partition-sum launch with the crossbar (bfir_run_device_quad)
    quad = pkg.Brutefir(L, P, rs, n_f, fmt, fmt, 48000, False, n_streams=S, xbar_inputs=n_in, xbar_outputs=n_out)
    assert quad.set_coeff(h, P) == 0
    quad.set_crossbar(gin, gout)
    out_q = [torch.zeros(n_o, dtype=tdt, device="cuda") for _ in range(nblk)]
    nq = nblk - nblk % 4
    for b in range(0, nq, 4):
        quad.run_device_quad(d_in[b:b + 4], out_q[b:b + 4])
    assert quad.sync() == 0 and quad.blockcounter() == nq
    for b in range(nq):
        assert rel_rms(out_q[b].cpu().numpy(), out_s[b].cpu().numpy()) < tol, b
    # the same through the stage pipeline (crossbar kernels on the side streams, block index from the host): staged
    # quads, a staged pair in between, a join
    stg = pkg.Brutefir(L, P, rs, n_f, fmt, fmt, 48000, False, n_streams=S, xbar_inputs=n_in, xbar_outputs=n_out, n_groups=1)
    assert stg.set_coeff(h, P) == 0
    stg.set_crossbar(gin, gout)
    out_t = [torch.zeros(n_o, dtype=tdt, device="cuda") for _ in range(nblk)]
    torch.cuda.synchronize()
    b = 0
    while b < nblk:
        if b + 4 <= nblk and b != 8:
            stg.run_device_quad(d_in[b:b + 4], out_t[b:b + 4], staged=True)
            b += 4
        else:
            stg.run_device_pair(d_in[b], d_in[b + 1], out_t[b], out_t[b + 1], pipelined="staged")
            b += 2
        if b == 12:
            stg.join()
    assert stg.sync() == 0 and stg.blockcounter() == nblk
    for b in range(nblk):
        assert rel_rms(out_t[b].cpu().numpy(), out_s[b].cpu().numpy()) < tol, b
