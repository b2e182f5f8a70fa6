"""GPU parity of the block loop (class brutefir, reference brutefir/brutefir.cpp:245-343) through
bfir_run: CUDA path vs the CPU oracle on the same seeded inputs, plus size-independent properties
(run() == direct linear convolution) at the BASELINE.json sizes."""
import numpy as np
import pytest

from conftest import white_noise, decay_filter, rel_rms, encode_raw, decode_raw, parity, SwapChain

pytestmark = pytest.mark.gpu

TOL = {4: 1e-5, 8: 1e-12}


def run_both(pkg, oracle, L, P, rs, C, in_fmt, out_fmt, n_blocks, dither=False, rate=44100, taps=None,
             coeff_blocks=None, scale=1.0, amp=1.0, seed=0):
    taps = L * P if taps is None else taps
    coeff_blocks = P if coeff_blocks is None else coeff_blocks
    g = pkg.Brutefir(L, P, rs, C, in_fmt, out_fmt, rate, dither)
    o = oracle.Engine(L, P, rs, C, in_fmt, out_fmt, rate, dither)
    h = [decay_filter(c, taps) for c in range(C)]
    assert g.set_coeff(h, coeff_blocks, scale) == 0
    assert o.set_coeff(h, coeff_blocks, scale) == 0
    assert g.is_initialized() and o.is_initialized()
    x = white_noise(seed, n_blocks * L, C) * amp
    outs_g, outs_o = [], []
    for b in range(n_blocks):
        raw = encode_raw(x[b * L:(b + 1) * L], in_fmt)
        rc_g, out_g = g.run(raw)
        rc_o, out_o = o.run(raw)
        assert rc_g == 0 and rc_o == 0
        outs_g.append(decode_raw(out_g, out_fmt, C))
        outs_o.append(decode_raw(out_o, out_fmt, C))
    return g, o, x, h, np.concatenate(outs_g), np.concatenate(outs_o)


@pytest.mark.parametrize("L,P,rs,C", [(64, 3, 4, 2), (256, 5, 8, 3), (1024, 4, 4, 8), (16, 2, 8, 1), (512, 1, 4, 2),
                                      (2048, 7, 8, 2), (128, 16, 4, 11)])
def test_run_matches_oracle_float_io(pkg, oracle, L, P, rs, C):
    fmt = pkg.FLOAT_LE if rs == 4 else pkg.FLOAT64_LE
    g, o, x, h, yg, yo = run_both(pkg, oracle, L, P, rs, C, fmt, fmt, 3 * P + 2, taps=L * P - 3)
    for c in range(C):
        assert rel_rms(yg[:, c], yo[:, c]) < TOL[rs]
    assert g.blockcounter() == o.blockcounter() == 3 * P + 2
    for c in range(C):
        a, b = g.overflow(c), o.overflow(c)
        # a sample within rounding distance of +-1.0 may fall on either side of the threshold
        assert abs(int(a.n_overflows) - int(b.n_overflows)) <= 2 and a.max == b.max
        assert abs(a.largest - b.largest) <= 1e-4 * max(1.0, b.largest)


def test_cfg0_stereo_float_vs_oracle_and_direct(pkg, oracle):
    """BASELINE configs[0]: stereo float, 65536 taps, L=4096, P=16"""
    from oracle import oracle_np
    L, P = 4096, 16
    g, o, x, h, yg, yo = run_both(pkg, oracle, L, P, 4, 2, pkg.FLOAT_LE, pkg.FLOAT_LE, 24)
    for c in range(2):
        assert rel_rms(yg[:, c], yo[:, c]) < 1e-5
        truth = oracle_np.direct_convolution(x[:, c].astype(np.float32), h[c].astype(np.float32), len(x))
        assert rel_rms(yg[:, c], truth) < 1e-5


def test_cfg1_double_vs_oracle_and_direct(pkg, oracle):
    """BASELINE configs[1]: 7.1 (8 ch) double, 262144 taps, L=8192, P=32"""
    from oracle import oracle_np
    L, P, C = 8192, 32, 8
    g, o, x, h, yg, yo = run_both(pkg, oracle, L, P, 8, C, pkg.FLOAT64_LE, pkg.FLOAT64_LE, 36)
    for c in range(C):
        assert rel_rms(yg[:, c], yo[:, c]) < 1e-12
        assert rel_rms(yg[:, c], oracle_np.direct_convolution(x[:, c], h[c], len(x))) < 1e-12


@pytest.mark.parametrize("L,rs", [(32768, 4), (16384, 8), (16384, 4), (8192, 8), (32768, 8)])
def test_largest_block_lengths_two_cta_transform(pkg, oracle, L, rs):
    """L = 32768 (float) / 16384 (double) only exist as two-CTA transforms (cfg4 uses L = 32768); L = 32768 in double
    precision (65536-point transforms) runs on FOUR CTAs per transform, the forward side as a thread-block cluster
    whose CTAs 1 and 3 exchange their sub-transforms through distributed shared memory"""
    fmt = pkg.FLOAT_LE if rs == 4 else pkg.FLOAT64_LE
    g, o, x, h, yg, yo = run_both(pkg, oracle, L, 3, rs, 2, fmt, fmt, 5)
    assert rel_rms(yg, yo) < TOL[rs]


def test_product_configuration_float_io_double_engine(pkg, oracle):
    """what foo_dsp_bfir constructs: FILTER_LEN 1024, REALSIZE 8, FLOAT_LE in/out (foo_dsp_bfir.cpp:279-286)"""
    g, o, x, h, yg, yo = run_both(pkg, oracle, 1024, 8, 8, 2, pkg.FLOAT_LE, pkg.FLOAT_LE, 20)
    assert rel_rms(yg, yo) < 1e-7     # output is rounded to float32 on both sides


@pytest.mark.parametrize("in_fmt", list(range(1, 12)))
def test_all_input_formats(pkg, oracle, in_fmt):
    g, o, x, h, yg, yo = run_both(pkg, oracle, 256, 3, 4, 2, in_fmt, pkg.FLOAT_LE, 8)
    assert rel_rms(yg, yo) < 1e-5


@pytest.mark.parametrize("rs", [4, 8])
@pytest.mark.parametrize("out_fmt", [1, 2, 3, 4, 5, 6, 7])
def test_integer_output_no_dither_within_1lsb(pkg, oracle, rs, out_fmt):
    in_fmt = pkg.FLOAT_LE
    g, o, x, h, yg, yo = run_both(pkg, oracle, 512, 4, rs, 2, in_fmt, out_fmt, 10, amp=1.3)
    bits = 8 * pkg.FORMAT_BYTES[out_fmt]
    if rs == 8 or bits <= 16:
        assert np.max(np.abs(yg - yo)) <= 1      # within 1 LSB with dither off (north star)
        assert np.mean(yg == yo) > 0.99
        lsb_tol = 1
    else:
        # a float32 engine carries 24 significant bits: for 24/32-bit output one LSB is below the
        # rounding noise of ANY float32 FFT (FFTW plans differ among themselves by more), so the
        # 1e-5 relative RMS gate applies instead
        assert rel_rms(yg, yo) < 1e-5
        lsb_tol = 2 ** (bits - 16)
    for c in range(2):
        a, b = g.overflow(c), o.overflow(c)
        assert abs(int(a.n_overflows) - int(b.n_overflows)) <= 2 and b.n_overflows > 0
        assert abs(a.intlargest - b.intlargest) <= lsb_tol


@pytest.mark.parametrize("rs", [4, 8])
def test_dithered_output_table_walk_and_lsb(pkg, oracle, rs):
    """dither on: the table walk (randtab_ptr per block) is bit-exact; samples stay within 1 LSB of the
    oracle except where the (different-rounding) convolution output flips a quantiser decision"""
    # rate 3300 -> table spacing 33000, 70 blocks: channel 0 walks [1, 17921], channel 1 [33001, 50921] --
    # clear of table indices 18310, 21146, 32686 where tab[n]-tab[n-1] == +255 makes the REFERENCE read one
    # element past its dither map (dither.cpp:77-78,160-161: its output there changes from run to run)
    L, P, C, rate = 256, 3, 2, 3300
    g, o, x, h, yg, yo = run_both(pkg, oracle, L, P, rs, C, pkg.FLOAT_LE, pkg.S16_LE, 70, dither=True, rate=rate)
    for c in range(C):
        assert g.dither_ptr(c) == o.dither_ptr(c)
    # the requantiser is a chaotic recurrence: once a rounding-level difference of the convolution output
    # flips one decision, the error-feedback states differ and the two 1-LSB noise sequences decorrelate
    # (bit-exactness on identical input is pinned by test_gpu_conv.py::test_cbuf2raw_dither_bit_exact)
    assert np.max(np.abs(yg - yo)) <= 3
    assert np.mean(np.abs(yg - yo)) < 1.0


def test_reset_keeps_buffers_and_restarts_counters(pkg, oracle):
    L, P, C = 128, 4, 2
    g = pkg.Brutefir(L, P, 4, C, pkg.FLOAT_LE, pkg.FLOAT_LE, 44100, False)
    o = oracle.Engine(L, P, 4, C, pkg.FLOAT_LE, pkg.FLOAT_LE, 44100, False)
    h = [decay_filter(c, L * P) for c in range(C)]
    g.set_coeff(h, P); o.set_coeff(h, P)
    x = white_noise(3, 12 * L, C)
    for b in range(12):
        if b == 6:
            g.reset(); o.reset()     # brutefir.cpp:347-367: counters only; stale time block remains
            assert g.blockcounter() == 0
        raw = encode_raw(x[b * L:(b + 1) * L], pkg.FLOAT_LE)
        _, yg = g.run(raw)
        _, yo = o.run(raw)
        assert rel_rms(yg.view(np.float32), yo.view(np.float32)) < 1e-5


def test_nonfinite_input_returns_minus_one_and_does_not_advance(pkg, oracle):
    L, P, C = 64, 2, 3
    g = pkg.Brutefir(L, P, 4, C, pkg.FLOAT_LE, pkg.FLOAT_LE, 44100, False)
    o = oracle.Engine(L, P, 4, C, pkg.FLOAT_LE, pkg.FLOAT_LE, 44100, False)
    h = [decay_filter(c, L * P) for c in range(C)]
    g.set_coeff(h, P); o.set_coeff(h, P)
    x = white_noise(5, L, C).astype(np.float32)
    assert g.run(x.view(np.uint8).ravel())[0] == 0 and o.run(x.view(np.uint8).ravel())[0] == 0
    bad = x.copy(); bad[3, 1] = np.nan
    assert g.run(bad.view(np.uint8).ravel())[0] == -1       # brutefir.cpp:316-321
    assert o.run(bad.view(np.uint8).ravel())[0] == -1
    assert g.blockcounter() == o.blockcounter() == 1        # :337-340 not reached


def test_set_coeff_errors_and_replacement(pkg, oracle):
    L, P, C = 64, 2, 2
    g = pkg.Brutefir(L, P, 4, C, pkg.FLOAT_LE, pkg.FLOAT_LE, 44100, False)
    assert not g.is_initialized()
    x = white_noise(1, L, C).astype(np.float32).view(np.uint8).ravel()
    with pytest.raises(pkg.BfirError):
        g.run(x)                                            # run before set_coeff
    h = [decay_filter(c, L * P) for c in range(C)]
    hbad = [h[0].copy(), h[1].copy()]
    hbad[1][5] = np.inf
    assert g.set_coeff(hbad, P) == -2 and not g.is_initialized()   # brutefir.cpp:219-224
    assert g.set_coeff(h, P) == 0 and g.is_initialized()
    y1 = g.run(x)[1].copy()
    assert g.set_coeff([2 * h[0], 2 * h[1]], P) == 0               # replacing filters keeps the delay line
    g2 = pkg.Brutefir(L, P, 4, C, pkg.FLOAT_LE, pkg.FLOAT_LE, 44100, False)
    g2.set_coeff(h, P, 2.0)
    assert rel_rms(g2.run(x)[1].view(np.float32), 2 * y1.view(np.float32)) < 1e-6


def test_invalid_parameters(pkg):
    for args in [(100, 2, 4, 2, 8, 8, 44100, False), (64, 2, 5, 2, 8, 8, 44100, False), (64, 0, 4, 2, 8, 8, 44100, False),
                 (64, 2, 4, 0, 8, 8, 44100, False), (64, 2, 4, 2, 0, 8, 44100, False), (64, 2, 4, 2, 8, 12, 44100, False),
                 (65536, 2, 4, 2, 8, 8, 44100, False), (65536, 2, 8, 2, 8, 8, 44100, False)]:
        with pytest.raises(pkg.BfirError) as e:
            pkg.Brutefir(*args)
        assert e.value.code == pkg.ERR_INVALID


def test_streams_batch_equals_separate_engines(pkg, oracle):
    """n_streams batching (cfg3 layout [stream][frame][channel]) == independent engines"""
    L, P, C, S = 256, 3, 2, 5
    g = pkg.Brutefir(L, P, 4, C, pkg.FLOAT_LE, pkg.FLOAT_LE, 44100, False, n_streams=S)
    h = [decay_filter(c, L * P) for c in range(C * S)]
    assert g.set_coeff(h, P) == 0
    singles = []
    for s in range(S):
        e = oracle.Engine(L, P, 4, C, pkg.FLOAT_LE, pkg.FLOAT_LE, 44100, False)
        e.set_coeff(h[s * C:(s + 1) * C], P)
        singles.append(e)
    x = white_noise(9, 6 * L, C * S).astype(np.float32)
    for b in range(6):
        blk = x[b * L:(b + 1) * L]                                  # [L, S*C]
        batched = np.ascontiguousarray(blk.reshape(L, S, C).transpose(1, 0, 2))   # [S, L, C]
        _, out = g.run(batched.view(np.uint8).ravel())
        out = out.view(np.float32).reshape(S, L, C)
        for s in range(S):
            _, ref = singles[s].run(np.ascontiguousarray(batched[s]).view(np.uint8).ravel())
            assert rel_rms(out[s], ref.view(np.float32).reshape(L, C)) < 1e-5


@pytest.mark.parametrize("out_fmt,dither", [(8, False), (2, True)])
def test_channel_groups_do_not_change_results(pkg, oracle, out_fmt, dither):
    """group pipelining (n_groups CUDA streams) is a scheduling choice only: bit-identical output"""
    import torch
    L, P, C, S = 256, 4, 2, 7
    h = [decay_filter(c, L * P) for c in range(C * S)]
    engines = [pkg.Brutefir(L, P, 4, C, pkg.FLOAT_LE, out_fmt, 2000, dither, n_streams=S, n_groups=g) for g in (1, 3, 8)]
    assert [e.get_groups() for e in engines] == [1, 3, 7]
    for e in engines:
        assert e.set_coeff(h, P) == 0
    x = white_noise(11, 9 * L, C * S).astype(np.float32)
    nb = pkg.FORMAT_BYTES[out_fmt]
    d_in = torch.empty(S * L * C, dtype=torch.float32, device="cuda")
    for b in range(9):
        blk = np.ascontiguousarray(x[b * L:(b + 1) * L].reshape(L, S, C).transpose(1, 0, 2))
        outs = []
        for i, e in enumerate(engines):
            if b % 2 == 0:
                rc, out = e.run(blk.view(np.uint8).ravel())
                assert rc == 0
            else:                                   # asynchronous device-buffer path
                d_in.copy_(torch.from_numpy(blk.ravel()))
                d_out = torch.zeros(S * L * C * nb, dtype=torch.uint8, device="cuda")
                torch.cuda.synchronize()
                e.run_device(d_in, d_out)
                assert e.sync() == 0
                out = d_out.cpu().numpy()
            outs.append(out.copy())
        assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2]), b
        if b == 4:                                  # switching the group count mid-stream keeps the state
            engines[1].set_groups(2)
    for e in engines:
        assert e.blockcounter() == 9


def test_partition_shards_sum_to_full_filter(pkg, oracle):
    """partition sharding: partial spectra of two shards summed == unsharded engine"""
    import torch
    L, P, C = 512, 8, 2
    h = [decay_filter(c, L * P) for c in range(C)]
    full = pkg.Brutefir(L, P, 4, C, pkg.FLOAT_LE, pkg.FLOAT_LE, 44100, False)
    a = pkg.Brutefir(L, P, 4, C, pkg.FLOAT_LE, pkg.FLOAT_LE, 44100, False, part_begin=0, part_count=3)
    b = pkg.Brutefir(L, P, 4, C, pkg.FLOAT_LE, pkg.FLOAT_LE, 44100, False, part_begin=3, part_count=5)
    for e in (full, a, b):
        assert e.set_coeff(h, P) == 0
    pa, n = a.acc_device_ptr()
    pb, _ = b.acc_device_ptr()
    ta, tb = pkg.as_torch(pa, n // 4, "<f4"), pkg.as_torch(pb, n // 4, "<f4")
    x = white_noise(2, 12 * L, C).astype(np.float32)
    d_in = torch.empty(L * C, dtype=torch.float32, device="cuda")
    outs = [torch.empty(L * C, dtype=torch.float32, device="cuda") for _ in range(3)]
    for blk in range(12):
        d_in.copy_(torch.from_numpy(x[blk * L:(blk + 1) * L].ravel()))
        torch.cuda.synchronize()
        full.run_device(d_in, outs[0])
        a.run_partial_device(d_in)
        b.run_partial_device(d_in)
        assert full.sync() == 0 and a.sync() == 0 and b.sync() == 0
        ta += tb                                   # what the NCCL reduce does across GPUs
        torch.cuda.synchronize()
        a.run_finish_device(outs[1])
        b.run_finish_device(outs[2])               # keeps shard b's block counter in step
        assert a.sync() == 0 and b.sync() == 0
        assert rel_rms(outs[1].cpu().numpy(), outs[0].cpu().numpy()) < 1e-6, blk


@pytest.mark.parametrize("rs,L,P,C", [(4, 256, 4, 2), (8, 128, 3, 2), (4, 4096, 2, 2)])
def test_crossfade_filter_swap_every_block(pkg, oracle, rs, L, P, C):
    """BASELINE configs[2]: a new coefficient set every block, output = crossfade_inplace(old, new).
    Oracle = the reference's entry points composed in run() order (conftest.SwapChain), with
    convolver_crossfade_inplace (fftw_convolver.cpp:276-321; float-branch algorithm for double)
    between the partition sums and the output stage; for the float build the same chain in double is the truth."""
    fmt = pkg.FLOAT_LE if rs == 4 else pkg.FLOAT64_LE
    dt = np.float32 if rs == 4 else np.float64
    nb = 2 * P + 3
    filters = [[decay_filter(100 * b + c, L * P).astype(dt) for c in range(C)] for b in range(nb + 1)]
    g = pkg.Brutefir(L, P, rs, C, fmt, fmt, 96000, False)
    assert g.set_coeff(filters[0], P) == 0
    ref = SwapChain(oracle.Convolver(L, rs, kind="port" if rs == 8 else None), L, P, C, dt)
    tru = SwapChain(oracle.Convolver(L, 8, kind="port"), L, P, C, np.float64) if rs == 4 else None
    H = [ref.spectra(fs) for fs in filters]
    H64 = [tru.spectra(fs) for fs in filters] if tru else None
    x = white_noise(33, nb * L, C).astype(dt)
    got, want, truth = [], [], []
    for t in range(nb):
        swap = t >= 1                                  # first block plain, then a swap on every block
        if swap:
            assert g.set_coeff_crossfade(filters[t], P) == 0
        blk = np.ascontiguousarray(x[t * L:(t + 1) * L])
        rc, out = g.run(blk.view(np.uint8).ravel())
        assert rc == 0
        got.append(out.view(dt).reshape(L, C).copy())
        want.append(ref.block(blk, H[t] if swap else H[0], H[t - 1] if swap else None))
        if tru:
            truth.append(tru.block(blk, H64[t] if swap else H64[0], H64[t - 1] if swap else None))
    got, want = np.concatenate(got), np.concatenate(want)
    truth = np.concatenate(truth) if tru else None
    for c in range(C):
        parity("engine_crossfade_swap/rs%d/L%d/P%d/ch%d" % (rs, L, P, c), got[:, c], want[:, c], TOL[rs],
               None if truth is None else truth[:, c])
    # a swap needs the same geometry and an initialised engine
    with pytest.raises(pkg.BfirError):
        g.set_coeff_crossfade(filters[0], P + 1)


def test_dense_filter_steady_state_512_partitions(pkg, oracle):
    """fp32 accumulation over 512 partitions in steady state (the sqrt(P) 2^-24 budget), slot
    arithmetic modulo P + 3 with every partition live: dense 524288-tap filters, L 1024, 530 blocks -- the last
    18 blocks have all 512 partitions contributing. Against the reference engine (float) and against direct linear
    convolution in float64. Also through the two-block entry point (pair kernel: contiguous partition runs)."""
    import torch
    from oracle import oracle_np
    L, P, C, nb = 1024, 512, 2, 530
    h = [decay_filter(40 + c, L * P).astype(np.float32) for c in range(C)]
    g = pkg.Brutefir(L, P, 4, C, pkg.FLOAT_LE, pkg.FLOAT_LE, 48000, False)
    g2 = pkg.Brutefir(L, P, 4, C, pkg.FLOAT_LE, pkg.FLOAT_LE, 48000, False)
    o = oracle.Engine(L, P, 4, C, pkg.FLOAT_LE, pkg.FLOAT_LE, 48000, False)
    for e in (g, g2, o):
        assert e.set_coeff(h, P) == 0
    x = white_noise(77, nb * L, C).astype(np.float32)
    yg, yo = np.empty_like(x), np.empty_like(x)
    for b in range(nb):
        raw = np.ascontiguousarray(x[b * L:(b + 1) * L]).view(np.uint8).ravel()
        rc, out = g.run(raw)
        rc_o, out_o = o.run(raw)
        assert rc == 0 and rc_o == 0
        yg[b * L:(b + 1) * L] = out.view(np.float32).reshape(L, C)
        yo[b * L:(b + 1) * L] = out_o.view(np.float32).reshape(L, C)
    # the same blocks two per call on device buffers (pairs engage once P blocks have been seen)
    d_x = torch.from_numpy(x).cuda()
    d_y = torch.empty_like(d_x)
    torch.cuda.synchronize()
    for b in range(0, nb, 2):
        g2.run_device_pair(d_x[b * L:(b + 1) * L], d_x[(b + 1) * L:(b + 2) * L], d_y[b * L:(b + 1) * L], d_y[(b + 1) * L:(b + 2) * L])
    assert g2.sync() == 0
    yp = d_y.cpu().numpy()
    steady = slice(P * L, nb * L)                      # blocks 512 .. 529: every partition has history
    for c in range(C):
        truth = oracle_np.direct_convolution(x[:, c], h[c], nb * L)
        parity("dense_P512/steady/ch%d" % c, yg[steady, c], yo[steady, c], 1e-5, truth[steady])
        parity("dense_P512/all_blocks/ch%d" % c, yg[:, c], yo[:, c], 1e-5, truth)
        parity("dense_P512/pairs_steady/ch%d" % c, yp[steady, c], yo[steady, c], 1e-5, truth[steady])
        assert rel_rms(yg[steady, c], truth[steady]) < 1e-5 and rel_rms(yp[steady, c], truth[steady]) < 1e-5


def test_cfg3_full_size_properties(pkg, oracle):
    """BASELINE configs[3] at full size on one GPU: 4096 independent stereo streams x 65536 taps
    (L 4096, P 16, float, distinct filter per channel; 8 GiB of spectra). 32 of the streams (64 channels) carry DENSE
    65536-tap filters and are run through the CPU oracle (one reference engine per stream); the other channels'
    filters are distinct gain * delay pairs, so their output must be the delayed, scaled input (exact overlap-save).
    Streams must not leak into each other, and the 4-group pipelined run must equal the serialised one."""
    import torch
    L, P, C, S = 4096, 16, 2, 4096
    Ct = C * S
    rng = np.random.default_rng(3)
    delays = rng.integers(0, L * P - 1, Ct)
    gains = rng.uniform(0.25, 1.0, Ct).astype(np.float32)
    dense_streams = sorted(int(s) for s in rng.choice(S, 32, replace=False))
    dense = {s * C + k: decay_filter(5000 + s * C + k, L * P).astype(np.float32) for s in dense_streams for k in range(C)}

    def filters():
        out = []
        for c in range(Ct):
            if c in dense:
                out.append(dense[c])
                continue
            v = np.zeros(L * P, dtype=np.float32)
            v[delays[c]] = gains[c]
            out.append(v)
        return out

    g = pkg.Brutefir(L, P, 4, C, pkg.FLOAT_LE, pkg.FLOAT_LE, 44100, False, n_streams=S, n_groups=1)
    assert g.set_coeff(filters(), P) == 0
    nb = 20
    gen = torch.Generator(device="cuda").manual_seed(5)
    x = [torch.rand(S, L, C, dtype=torch.float32, device="cuda", generator=gen) * 2 - 1 for _ in range(nb)]
    y = [torch.empty(S, L, C, dtype=torch.float32, device="cuda") for _ in range(nb)]
    torch.cuda.synchronize()                           # the engine runs on its own stream
    for b in range(nb):
        g.run_device(x[b], y[b])
    assert g.sync() == 0
    X = torch.cat(x, dim=1)                             # [S, nb*L, C]
    Y = torch.cat(y, dim=1)
    sparse = [c for c in range(Ct) if c not in dense]
    for c in rng.choice(sparse, 64, replace=False):
        s, k, d = int(c) // C, int(c) % C, int(delays[c])
        want = torch.zeros(nb * L, device="cuda")
        want[d:] = X[s, : nb * L - d, k] * float(gains[c])
        err = torch.sqrt(torch.mean((Y[s, :, k] - want) ** 2) / torch.mean(want ** 2)).item()
        assert err < 1e-5, (c, err)
    # the dense-filter streams against the reference engine on the same samples
    worst = 0.0
    for s in dense_streams:
        o = oracle.Engine(L, P, 4, C, pkg.FLOAT_LE, pkg.FLOAT_LE, 44100, False)
        assert o.set_coeff([dense[s * C + k] for k in range(C)], P) == 0
        xs = X[s].cpu().numpy()
        ys = Y[s].cpu().numpy()
        ref = np.empty_like(ys)
        for b in range(nb):
            rc_o, out_o = o.run(np.ascontiguousarray(xs[b * L:(b + 1) * L]).view(np.uint8).ravel())
            assert rc_o == 0
            ref[b * L:(b + 1) * L] = out_o.view(np.float32).reshape(L, C)
        for k in range(C):
            worst = max(worst, rel_rms(ys[:, k], ref[:, k]))
    parity("cfg3_full_size/dense_streams_worst_channel", [worst], [0.0], 1e-5)
    # pipelined (4 channel groups) == serialised, bit for bit, at full size
    g2 = pkg.Brutefir(L, P, 4, C, pkg.FLOAT_LE, pkg.FLOAT_LE, 44100, False, n_streams=S, n_groups=4)
    assert g2.set_coeff(filters(), P) == 0
    y2 = torch.empty(S, L, C, dtype=torch.float32, device="cuda")
    for b in range(nb):
        g2.run_device(x[b], y2)
        assert g2.sync() == 0
        assert torch.equal(y2, y[b]), b


@pytest.mark.parametrize("P,cb,taps", [(6, 3, 3 * 128), (6, 6, 200), (4, 4, 1), (5, 2, 2 * 128 - 1)])
def test_ragged_coefficient_geometry(pkg, oracle, P, cb, taps):
    """coefficient sets shorter than the engine (coeff_blocks < filter_blocks, brutefir.cpp:292) and filters
    that end inside a partition (coeff::preprocess_coeff zero-padding, coeff.cpp:317-341)"""
    L, C = 128, 2
    g = pkg.Brutefir(L, P, 8, C, pkg.FLOAT64_LE, pkg.FLOAT64_LE, 44100, False)
    o = oracle.Engine(L, P, 8, C, pkg.FLOAT64_LE, pkg.FLOAT64_LE, 44100, False)
    h = [decay_filter(c, taps) for c in range(C)]
    assert g.set_coeff(h, cb, 0.7) == 0 and o.set_coeff(h, cb, 0.7) == 0
    x = white_noise(17, 3 * P * L, C)
    for b in range(3 * P):
        raw = np.ascontiguousarray(x[b * L:(b + 1) * L]).view(np.uint8).ravel()
        yg, yo = g.run(raw)[1].view(np.float64), o.run(raw)[1].view(np.float64)
        assert rel_rms(yg, yo) < 1e-12, b


def test_fewer_coefficient_sets_than_channels(pkg, oracle):
    """set_coeff with n_coeffs < channels: the reference leaves coeffs[n].data NULL and run() would crash
    (brutefir.cpp:213-216, 283-297); the build convolves the channels that have a filter and emits silence
    for the others. n_coeffs > channels is clamped like the reference (:191-194)."""
    L, P, C = 256, 3, 4
    g = pkg.Brutefir(L, P, 4, C, pkg.FLOAT_LE, pkg.FLOAT_LE, 44100, False)
    o = oracle.Engine(L, P, 4, 2, pkg.FLOAT_LE, pkg.FLOAT_LE, 44100, False)
    h = [decay_filter(c, L * P) for c in range(6)]
    assert g.set_coeff(h[:2], P) == 0 and o.set_coeff(h[:2], P) == 0
    x = white_noise(2, 5 * L, C).astype(np.float32)
    for b in range(5):
        blk = np.ascontiguousarray(x[b * L:(b + 1) * L])
        y = g.run(blk.view(np.uint8).ravel())[1].view(np.float32).reshape(L, C)
        ref = o.run(np.ascontiguousarray(blk[:, :2]).view(np.uint8).ravel())[1].view(np.float32).reshape(L, 2)
        assert rel_rms(y[:, :2], ref) < 1e-5 and np.all(y[:, 2:] == 0)
    assert g.set_coeff(h, P) == 0          # six arrays for four channels: clamped
    assert g.run(np.ascontiguousarray(x[:L]).view(np.uint8).ravel())[0] == 0


def test_check_overflows_prints_peaks(pkg):
    """brutefir::check_overflows / print_overflows (brutefir.cpp:371-388, 585-629) through the print callback"""
    msgs = []
    cb = pkg.PRINT_CB(lambda m: msgs.append(m.decode()))
    pkg.load_library().bfir_set_print_callback(cb)
    try:
        L, P, C = 128, 2, 2
        g = pkg.Brutefir(L, P, 4, C, pkg.FLOAT_LE, pkg.S16_LE, 44100, False)
        g.set_coeff([np.array([4.0]), np.array([0.25])], P)
        x = (white_noise(1, L, C) * 0.9).astype(np.float32)
        g.run(x.view(np.uint8).ravel())
        assert g.check_overflows() == 1              # channel 0 clips (gain 4), channel 1 does not
        assert len(msgs) == 2 and msgs[0].startswith("peak: 0/") and msgs[1].startswith("peak: 1/0/")
        assert g.overflow(0).n_overflows > 0 and g.overflow(1).n_overflows == 0
        assert g.check_overflows() == 0              # unchanged since the last call: silent
    finally:
        pkg.load_library().bfir_set_print_callback(pkg.PRINT_CB(0))


@pytest.mark.parametrize("groups", [1, 2, 4])
@pytest.mark.parametrize("out_fmt,dither", [(8, False), (2, True)])
def test_run_async_equals_run(pkg, groups, out_fmt, dither, monkeypatch):
    """bfir_run_async / bfir_wait: the pipelined variant is a scheduling choice only -- bit-identical
    output to the synchronous bfir_run (one-kernel partition sum on both sides: BFIR_LOOKAHEAD=0), with
    several blocks in flight and the ticket ring wrapping"""
    import torch
    L, P, C, S = 256, 4, 2, 5
    h = [decay_filter(c, L * P) for c in range(C * S)]
    monkeypatch.setenv("BFIR_LOOKAHEAD", "0")
    sync_e = pkg.Brutefir(L, P, 4, C, pkg.FLOAT_LE, out_fmt, 2000, dither, n_streams=S, n_groups=1)
    async_e = pkg.Brutefir(L, P, 4, C, pkg.FLOAT_LE, out_fmt, 2000, dither, n_streams=S, n_groups=groups)
    assert sync_e.set_coeff(h, P) == 0 and async_e.set_coeff(h, P) == 0
    nblk, nb = 27, pkg.FORMAT_BYTES[out_fmt]
    x = white_noise(5, nblk * L, C * S).astype(np.float32)
    blocks = [np.ascontiguousarray(x[b * L:(b + 1) * L].reshape(L, S, C).transpose(1, 0, 2)).ravel() for b in range(nblk)]
    want = []
    for b in range(nblk):
        rc, out = sync_e.run(blocks[b].view(np.uint8))
        assert rc == 0
        want.append(out.copy())
    pin_in = [torch.from_numpy(blk).pin_memory() for blk in blocks]
    pin_out = [torch.zeros(S * L * C * nb, dtype=torch.uint8).pin_memory() for _ in range(nblk)]
    tickets = []
    for b in range(nblk):
        if b == 13:                                # a synchronous call in between joins the queued steps first
            rc, out = async_e.run(blocks[b].view(np.uint8))
            assert rc == 0 and np.array_equal(out, want[b])
            tickets.append(None)
            continue
        tickets.append(async_e.run_async(pin_in[b].numpy(), pin_out[b].numpy()))
        if b >= 3 and b % 3 == 0 and tickets[b - 3] is not None:   # wait with three steps still queued behind
            assert async_e.wait(tickets[b - 3]) == 0
            assert np.array_equal(pin_out[b - 3].numpy(), want[b - 3])
    assert async_e.wait(tickets[-1]) == 0
    for b in range(nblk):
        if tickets[b] is not None:
            assert np.array_equal(pin_out[b].numpy(), want[b]), b
    assert async_e.blockcounter() == sync_e.blockcounter() == nblk
    for c in range(C * S):
        a, w = async_e.overflow(c), sync_e.overflow(c)
        assert (a.n_overflows, a.largest) == (w.n_overflows, w.largest)
    with pytest.raises(pkg.BfirError):
        async_e.wait(10 ** 6)                      # a ticket that was never handed out


def test_run_async_reports_nonfinite_at_wait(pkg):
    import torch
    L, P, C = 128, 2, 2
    e = pkg.Brutefir(L, P, 8, C, pkg.FLOAT64_LE, pkg.FLOAT64_LE, 2000, False, n_streams=4, n_groups=2)
    assert e.set_coeff([decay_filter(c, L * P) for c in range(4 * C)], P) == 0
    good = torch.from_numpy(white_noise(1, L, 4 * C).ravel().copy()).pin_memory()
    bad = good.clone().pin_memory()
    bad[5] = float("nan")
    outs = [torch.zeros(4 * L * C, dtype=torch.float64).pin_memory() for _ in range(3)]
    t0 = e.run_async(good.numpy(), outs[0].numpy())
    assert e.wait(t0) == 0
    t1 = e.run_async(bad.numpy(), outs[1].numpy())
    t2 = e.run_async(good.numpy(), outs[2].numpy())
    assert e.wait(t2) == -1                         # raised by the block behind t1, reported at the wait
    assert e.wait(t1) == 0 and e.sync() == 0        # reported once


def test_run_device_pipelined_equals_run_device(pkg):
    """bfir_run_device_pipelined + bfir_join: no join between blocks, same bits as bfir_run_device"""
    import torch
    L, P, C, S = 512, 3, 2, 6
    h = [decay_filter(c, L * P) for c in range(C * S)]
    a = pkg.Brutefir(L, P, 8, C, pkg.FLOAT64_LE, pkg.FLOAT64_LE, 2000, False, n_streams=S, n_groups=1)
    b = pkg.Brutefir(L, P, 8, C, pkg.FLOAT64_LE, pkg.FLOAT64_LE, 2000, False, n_streams=S, n_groups=3)
    assert a.set_coeff(h, P) == 0 and b.set_coeff(h, P) == 0
    st = torch.cuda.Stream()
    b.set_stream(st.cuda_stream)
    nblk = 11
    d_in = [torch.from_numpy(white_noise(40 + k, L, C * S).ravel().copy()).cuda() for k in range(nblk)]
    out_a = [torch.zeros(S * L * C, dtype=torch.float64, device="cuda") for _ in range(nblk)]
    out_b = [torch.zeros(S * L * C, dtype=torch.float64, device="cuda") for _ in range(nblk)]
    torch.cuda.synchronize()
    for k in range(nblk):
        a.run_device(d_in[k], out_a[k])
        b.run_device_pipelined(d_in[k], out_b[k])
    b.join()                                          # stream-ordered: work queued on `st` after this sees every block
    with torch.cuda.stream(st):
        total = torch.stack(out_b).sum()
    assert a.sync() == 0 and b.sync() == 0
    for k in range(nblk):
        assert torch.equal(out_a[k], out_b[k]), k
    assert float(total) == float(torch.stack(out_a).sum())
    assert b.blockcounter() == nblk


@pytest.mark.parametrize("rs,groups", [(4, 1), (8, 1), (4, 3)])
def test_lookahead_partition_sum_equals_plain_path(pkg, oracle, rs, groups, monkeypatch):
    """bfir_run computes partitions 1..P-1 of the next block ahead of time and folds partition 0 into the inverse
    transform's load phase; BFIR_LOOKAHEAD=0 keeps the one-kernel partition sum. Same result to rounding, through
    a filter replacement, a reset and a NaN abort (each of which must drop the look-ahead result)."""
    L, P, C, S = 256, 5, 2, 3
    fmt = pkg.FLOAT_LE if rs == 4 else pkg.FLOAT64_LE
    dt = np.float32 if rs == 4 else np.float64
    monkeypatch.setenv("BFIR_LOOKAHEAD", "0")
    plain = pkg.Brutefir(L, P, rs, C, fmt, fmt, 2000, False, n_streams=S, n_groups=groups)
    monkeypatch.setenv("BFIR_LOOKAHEAD", "1")
    ahead = pkg.Brutefir(L, P, rs, C, fmt, fmt, 2000, False, n_streams=S, n_groups=groups)
    orc = [oracle.Engine(L, P, rs, C, fmt, fmt, 2000, False) for _ in range(S)]
    h1 = [decay_filter(c, L * P) for c in range(C * S)]
    h2 = [decay_filter(100 + c, L * 3) for c in range(C * S)]
    for e in (plain, ahead):
        assert e.set_coeff(h1, P) == 0
    for s_, o in enumerate(orc):
        assert o.set_coeff(h1[s_ * C:(s_ + 1) * C], P) == 0
    x = white_noise(77, 30 * L, C * S).astype(dt)
    tol = 2e-6 if rs == 4 else 1e-13
    for b in range(30):
        blk = np.ascontiguousarray(x[b * L:(b + 1) * L].reshape(L, S, C).transpose(1, 0, 2))
        if b == 9:                                   # replace the filters (shorter: 3 partitions loaded)
            for e in (plain, ahead):
                assert e.set_coeff(h2, 3) == 0
            for s_, o in enumerate(orc):
                assert o.set_coeff(h2[s_ * C:(s_ + 1) * C], 3) == 0
        if b == 15:
            for e in (plain, ahead):
                e.reset()
            for o in orc:
                o.reset()
        if b == 21:                                  # NaN block: -1 from both, counters stay, the stream goes on
            bad = blk.copy()
            bad[0, 3, 1] = np.nan
            assert plain.run(bad.view(np.uint8).ravel())[0] == -1 and ahead.run(bad.view(np.uint8).ravel())[0] == -1
            for s_, o in enumerate(orc):
                rc_o, _ = o.run(np.ascontiguousarray(bad[s_]).view(np.uint8).ravel())
                assert rc_o == (-1 if s_ == 0 else 0)
            continue
        rc_p, out_p = plain.run(blk.view(np.uint8).ravel())
        rc_a, out_a = ahead.run(blk.view(np.uint8).ravel())
        assert rc_p == 0 and rc_a == 0
        yp, ya = out_p.view(dt), out_a.view(dt)
        if b < 21:
            assert rel_rms(ya, yp) < tol, b
            for s_, o in enumerate(orc):
                rc_o, out_o = o.run(np.ascontiguousarray(blk[s_]).view(np.uint8).ravel())
                assert rc_o == 0
                assert rel_rms(ya.reshape(S, L * C)[s_], out_o.view(dt)) < (1e-5 if rs == 4 else 1e-12), (b, s_)
        else:                                        # after the NaN block both engines carry the same NaN history
            assert np.array_equal(np.isnan(ya), np.isnan(yp))
            m = ~np.isnan(yp)
            assert rel_rms(ya[m], yp[m]) < tol if m.any() else True
    assert plain.blockcounter() == ahead.blockcounter()


@pytest.mark.parametrize("L,S", [(256, 40), (4096, 36), (64, 70)])
def test_double_precision_batches_above_64_buffers(pkg, oracle, L, S):
    """More than 64 double-precision transforms per launch take the other kernel variants (16 points per thread on the
    inverse side, 8 on the forward side up to 4096 points per CTA; below that count both sides use 8): every variant
    against the oracle, a few streams of the batch each."""
    P, C = 3, 2
    g = pkg.Brutefir(L, P, 8, C, pkg.FLOAT64_LE, pkg.FLOAT64_LE, 48000, False, n_streams=S)
    h = [decay_filter(c % 7, L * P) * (1.0 + 0.01 * c) for c in range(C * S)]
    assert g.set_coeff(h, P) == 0
    probe = [0, S // 2, S - 1]
    singles = {}
    for s in probe:
        e = oracle.Engine(L, P, 8, C, pkg.FLOAT64_LE, pkg.FLOAT64_LE, 48000, False)
        assert e.set_coeff(h[s * C:(s + 1) * C], P) == 0
        singles[s] = e
    x = white_noise(19, 5 * L, C * S)
    for b in range(5):
        batched = np.ascontiguousarray(x[b * L:(b + 1) * L].reshape(L, S, C).transpose(1, 0, 2))
        rc, out = g.run(batched.view(np.uint8).ravel())
        assert rc == 0
        out = out.view(np.float64).reshape(S, L, C)
        for s in probe:
            _, ref = singles[s].run(np.ascontiguousarray(batched[s]).view(np.uint8).ravel())
            assert rel_rms(out[s], ref.view(np.float64).reshape(L, C)) < 1e-12, (b, s)


@pytest.mark.parametrize("rs,groups,out_fmt,dither", [(4, 1, 8, False), (8, 2, 10, False), (4, 3, 2, True), (8, 1, 4, False), (4, 3, 8, False), (4, 1, 2, True), (4, 3, 2, False), (8, 1, 2, True)])
def test_block_pairs_equal_single_blocks(pkg, oracle, rs, groups, out_fmt, dither):
    """bfir_run_device_pair / bfir_run_async_pair: two blocks per partition-sum launch. Same output as block by block
    up to the summation order; the first filter_blocks blocks (delay line still filling) take the one-by-one path."""
    import torch
    L, P, C, S = 256, 6, 2, 3
    in_fmt = pkg.FLOAT_LE if rs == 4 else pkg.FLOAT64_LE
    dt, tdt = (np.float32, torch.float32) if rs == 4 else (np.float64, torch.float64)
    nb = pkg.FORMAT_BYTES[out_fmt]
    h = [decay_filter(c, L * P - 7) for c in range(C * S)]
    single = pkg.Brutefir(L, P, rs, C, in_fmt, out_fmt, 2000, dither, n_streams=S, n_groups=1)
    pair_d = pkg.Brutefir(L, P, rs, C, in_fmt, out_fmt, 2000, dither, n_streams=S, n_groups=groups)
    pair_h = pkg.Brutefir(L, P, rs, C, in_fmt, out_fmt, 2000, dither, n_streams=S, n_groups=groups)
    for e in (single, pair_d, pair_h):
        assert e.set_coeff(h, P) == 0
    nblk = 20
    x = white_noise(91, nblk * L, C * S).astype(dt)
    blocks = [np.ascontiguousarray(x[b * L:(b + 1) * L].reshape(L, S, C).transpose(1, 0, 2)).ravel() for b in range(nblk)]
    d_in = [torch.from_numpy(b).cuda() for b in blocks]
    pin_in = [torch.from_numpy(b).pin_memory() for b in blocks]
    n_out = S * L * C * nb
    out_s = [torch.zeros(n_out, dtype=torch.uint8, device="cuda") for _ in range(nblk)]
    out_p = [torch.zeros(n_out, dtype=torch.uint8, device="cuda") for _ in range(nblk)]
    out_h = [torch.zeros(n_out, dtype=torch.uint8).pin_memory() for _ in range(nblk)]
    torch.cuda.synchronize()
    for b in range(nblk):
        single.run_device(d_in[b], out_s[b])
    tickets = []
    for b in range(0, nblk, 2):
        pair_d.run_device_pair(d_in[b], d_in[b + 1], out_p[b], out_p[b + 1], pipelined=(b % 4 == 0))
        tickets.append(pair_h.run_async_pair(pin_in[b].numpy(), pin_in[b + 1].numpy(), out_h[b].numpy(), out_h[b + 1].numpy()))
    assert single.sync() == 0 and pair_d.sync() == 0 and pair_h.wait(tickets[-1]) == 0
    assert single.blockcounter() == pair_d.blockcounter() == pair_h.blockcounter() == nblk

    for b in range(nblk):
        a = out_s[b].cpu().numpy()
        for which, other in (("device pair", out_p[b].cpu().numpy()), ("host pair", out_h[b].numpy())):
            if out_fmt in (8, 10):                       # float output: rounding-level difference only
                fa, fo = a.view(dt if out_fmt == (8 if rs == 4 else 10) else np.float32), other.view(dt if out_fmt == (8 if rs == 4 else 10) else np.float32)
                assert rel_rms(fo, fa) < (2e-6 if rs == 4 else 1e-13), (which, b)
            else:                                        # integer output (decode_raw gives LSBs): one LSB apart at most
                ia = decode_raw(a, out_fmt, C).ravel().astype(np.float64)
                io = decode_raw(other, out_fmt, C).ravel().astype(np.float64)
                dd = np.abs(ia - io)
                if rs == 4 and nb > 2:                   # float32 engine, 24/32-bit output: one LSB is below float32 resolution
                    assert rel_rms(io, ia) < 2e-6, (which, b)
                else:                                    # the dither feedback carries a flipped LSB on for a few samples
                    assert dd.max() <= (4 if dither else 1), (which, b, dd.max(), dd.mean())
                    assert dd.mean() < 0.5
        if b < P and not dither:                         # the fall-back path is the single-block path itself
            assert np.array_equal(a, out_p[b].cpu().numpy()), b


def test_cfg1_block_pairs_vs_oracle(pkg, oracle):
    """BASELINE configs[1] geometry (8 ch, double, L 8192, P 32) through the two-block entry point, straight against
    the oracle: the pair kernel must meet the same 1e-12 bar as the one-block path."""
    import torch
    L, P, C = 8192, 32, 8
    g = pkg.Brutefir(L, P, 8, C, pkg.FLOAT64_LE, pkg.FLOAT64_LE, 48000, False)
    o = oracle.Engine(L, P, 8, C, pkg.FLOAT64_LE, pkg.FLOAT64_LE, 48000, False)
    h = [decay_filter(c, L * P) for c in range(C)]
    assert g.set_coeff(h, P) == 0 and o.set_coeff(h, P) == 0
    nblk = P + 6
    x = white_noise(23, nblk * L, C)
    d_out = [torch.zeros(L * C, dtype=torch.float64, device="cuda") for _ in range(2)]
    for b in range(0, nblk, 2):
        blk = [np.ascontiguousarray(x[(b + k) * L:(b + k + 1) * L]) for k in range(2)]
        d_in = [torch.from_numpy(v.ravel()).cuda() for v in blk]
        torch.cuda.synchronize()
        g.run_device_pair(d_in[0], d_in[1], d_out[0], d_out[1])
        assert g.sync() == 0
        for k in range(2):
            rc, ref = o.run(blk[k].view(np.uint8).ravel())
            assert rc == 0
            if b + k >= P - 2:                      # the last one-by-one blocks and every pair
                got = d_out[k].cpu().numpy().reshape(L, C)
                want = ref.view(np.float64).reshape(L, C)
                for c in range(C):
                    assert rel_rms(got[:, c], want[:, c]) < 1e-12, (b + k, c)
    assert g.blockcounter() == o.blockcounter() == nblk


@pytest.mark.parametrize("rs,out_fmt,dither", [(8, 10, False), (4, 8, False), (4, 2, True)])
def test_stage_pipeline_equals_single_blocks(pkg, rs, out_fmt, dither):
    """bfir_run_device_pair(BFIR_PAIR_STAGED) on a one-group engine runs the stage pipeline (forward transforms of the next
    pair under the pair sum, inverse transforms of the previous pair behind it, block index from the host): same
    output as block by block, through mode changes (join, synchronous run, plain pair) in between."""
    import torch
    L, P, C, S = 512, 5, 2, 4
    in_fmt = pkg.FLOAT_LE if rs == 4 else pkg.FLOAT64_LE
    dt = np.float32 if rs == 4 else np.float64
    nb = pkg.FORMAT_BYTES[out_fmt]
    h = [decay_filter(c, L * P) for c in range(C * S)]
    single = pkg.Brutefir(L, P, rs, C, in_fmt, out_fmt, 2000, dither, n_streams=S, n_groups=1)
    staged = pkg.Brutefir(L, P, rs, C, in_fmt, out_fmt, 2000, dither, n_streams=S, n_groups=1)
    assert single.set_coeff(h, P) == 0 and staged.set_coeff(h, P) == 0
    nblk = 40
    x = white_noise(33, nblk * L, C * S).astype(dt)
    blocks = [np.ascontiguousarray(x[b * L:(b + 1) * L].reshape(L, S, C).transpose(1, 0, 2)).ravel() for b in range(nblk)]
    d_in = [torch.from_numpy(b).cuda() for b in blocks]
    n_out = S * L * C * nb
    out_s = [torch.zeros(n_out, dtype=torch.uint8, device="cuda") for _ in range(nblk)]
    out_t = [torch.zeros(n_out, dtype=torch.uint8, device="cuda") for _ in range(nblk)]
    torch.cuda.synchronize()
    for b in range(nblk):
        single.run_device(d_in[b], out_s[b])
    b = 0
    while b < nblk:
        if b == 20:                                  # a plain pair and a synchronous block in the middle
            staged.run_device_pair(d_in[b], d_in[b + 1], out_t[b], out_t[b + 1], pipelined=False)
            b += 2
            rc, o = staged.run(blocks[b].view(np.uint8))
            assert rc == 0
            out_t[b].copy_(torch.from_numpy(o))
            b += 1
            staged.run_device(d_in[b], out_t[b])
            b += 1
            continue
        staged.run_device_pair(d_in[b], d_in[b + 1], out_t[b], out_t[b + 1], pipelined="staged" if b != 30 else True)
        if b == 12:
            staged.join()
        b += 2
    assert single.sync() == 0 and staged.sync() == 0
    assert single.blockcounter() == staged.blockcounter() == nblk
    for b in range(nblk):
        a, t = out_s[b].cpu().numpy(), out_t[b].cpu().numpy()
        if out_fmt in (8, 10):
            assert rel_rms(t.view(dt), a.view(dt)) < (2e-6 if rs == 4 else 1e-13), b
        else:
            dd = np.abs(decode_raw(a, out_fmt, C).ravel().astype(np.float64) - decode_raw(t, out_fmt, C).ravel())
            assert dd.max() <= 4 and dd.mean() < 0.5, (b, dd.max(), dd.mean())


@pytest.mark.parametrize("rs,out_fmt,P", [(8, 10, 5), (4, 8, 5), (8, 10, 2), (4, 2, 7), (8, 4, 3)])
def test_stage_pipeline_quads_equal_single_blocks(pkg, rs, out_fmt, P):
    """bfir_run_device_quad_staged: four blocks per call through the stage pipeline (four forward transforms on the
    forward stream under the previous call's partition sum, ONE four-block partition-sum launch -- both precisions
    -- and four inverse transforms behind it), mixed with staged pairs, a join, a plain quad and single blocks: same
    output as block by block. The delay line has P + 7 slots so that the next call's transforms never touch a slot
    the running sum still reads; P = 2 makes every slot turn over within two calls."""
    import torch
    L, C, S = 512, 2, 4
    in_fmt = pkg.FLOAT_LE if rs == 4 else pkg.FLOAT64_LE
    dt = np.float32 if rs == 4 else np.float64
    nb = pkg.FORMAT_BYTES[out_fmt]
    h = [decay_filter(c, L * P) for c in range(C * S)]
    single = pkg.Brutefir(L, P, rs, C, in_fmt, out_fmt, 2000, False, n_streams=S, n_groups=1)
    staged = pkg.Brutefir(L, P, rs, C, in_fmt, out_fmt, 2000, False, n_streams=S, n_groups=1)
    assert single.set_coeff(h, P) == 0 and staged.set_coeff(h, P) == 0
    nblk = 64
    x = white_noise(35, nblk * L, C * S).astype(dt)
    blocks = [np.ascontiguousarray(x[b * L:(b + 1) * L].reshape(L, S, C).transpose(1, 0, 2)).ravel() for b in range(nblk)]
    d_in = [torch.from_numpy(b).cuda() for b in blocks]
    n_out = S * L * C * nb
    out_s = [torch.zeros(n_out, dtype=torch.uint8, device="cuda") for _ in range(nblk)]
    out_t = [torch.zeros(n_out, dtype=torch.uint8, device="cuda") for _ in range(nblk)]
    torch.cuda.synchronize()
    for b in range(nblk):
        single.run_device(d_in[b], out_s[b])
    b = 0
    while b < nblk:
        if b in (24, 26):                            # staged pairs between staged quads
            staged.run_device_pair(d_in[b], d_in[b + 1], out_t[b], out_t[b + 1], pipelined="staged")
            b += 2
            continue
        if b == 40:                                  # a plain (joined) quad, then single blocks
            staged.run_device_quad(d_in[b:b + 4], out_t[b:b + 4])
            b += 4
            for _ in range(4):
                staged.run_device(d_in[b], out_t[b])
                b += 1
            continue
        staged.run_device_quad(d_in[b:b + 4], out_t[b:b + 4], staged=True)
        if b == 12:
            staged.join()
        b += 4
    assert single.sync() == 0 and staged.sync() == 0
    assert single.blockcounter() == staged.blockcounter() == nblk
    for b in range(nblk):
        a, t = out_s[b].cpu().numpy(), out_t[b].cpu().numpy()
        if out_fmt in (8, 10):
            assert rel_rms(t.view(dt), a.view(dt)) < (2e-6 if rs == 4 else 1e-13), b
        else:
            dd = np.abs(decode_raw(a, out_fmt, C).ravel().astype(np.float64) - decode_raw(t, out_fmt, C).ravel())
            assert dd.max() <= 1, (b, dd.max())      # S16 from the float engine, S24 from the double one: within 1 LSB


@pytest.mark.parametrize("rs,out_fmt,P,S,C,xb,cb", [(8, 10, 5, 20, 4, 0, 5), (4, 8, 5, 40, 4, 0, 5), (8, 10, 2, 20, 4, 0, 2), (4, 2, 9, 40, 4, 0, 9),
                                                    (8, 4, 3, 24, 4, 0, 3), (4, 8, 4, 8, 32, 32, 4), (4, 8, 6, 2, 2, 0, 6),
                                                    (8, 10, 9, 20, 4, 0, 6), (4, 8, 11, 40, 4, 0, 1)])
def test_eight_blocks_per_call_equal_single_blocks(pkg, rs, out_fmt, P, S, C, xb, cb):
    """bfir_run_device_oct: eight blocks per call with ONE partition-sum launch (partition_mac_oct_kernel: a thread owns
    8 reals of an ORD group in single precision, 4 in double precision, circular window of eight delay-line spectra),
    joined and staged, mixed with four- and two-block calls and single blocks, with a crossbar, with integer output;
    one case is too small for the one-slice kernel and falls back to two four-block calls. P = 2 turns every delay-line
    slot over within one call (P + 15 slots); cb < P: filters with fewer coefficient partitions than the engine has
    (brutefir.cpp:292). Same output as block by block."""
    import torch
    L = 2048
    in_fmt = pkg.FLOAT_LE if rs == 4 else pkg.FLOAT64_LE
    dt = np.float32 if rs == 4 else np.float64
    nb = pkg.FORMAT_BYTES[out_fmt]
    n_in = xb or C
    n_out = xb or C
    h = [decay_filter(c % 7, L * cb) * (1 + 0.01 * c) for c in range(C * S)]
    kw = dict(n_streams=S, n_groups=1)
    if xb:
        kw.update(xbar_inputs=xb, xbar_outputs=xb)
    single = pkg.Brutefir(L, P, rs, C, in_fmt, out_fmt, 2000, False, **kw)
    octs = pkg.Brutefir(L, P, rs, C, in_fmt, out_fmt, 2000, False, **kw)
    assert single.set_coeff(h, cb) == 0 and octs.set_coeff(h, cb) == 0
    if xb:
        rng = np.random.default_rng(3)
        gin, gout = rng.standard_normal((C, xb)) / np.sqrt(xb), rng.standard_normal((xb, C)) / np.sqrt(C)
        single.set_crossbar(gin, gout)
        octs.set_crossbar(gin, gout)
    nblk = 8 * 6 + 4 + 2 + 2
    gen = torch.Generator(device="cuda").manual_seed(7)
    tdt = torch.float32 if rs == 4 else torch.float64
    d_in = [torch.rand(S * L * n_in, dtype=tdt, device="cuda", generator=gen) * 2 - 1 for _ in range(nblk)]
    n_o = S * L * n_out * nb
    out_s = [torch.zeros(n_o, dtype=torch.uint8, device="cuda") for _ in range(nblk)]
    out_t = [torch.zeros(n_o, dtype=torch.uint8, device="cuda") for _ in range(nblk)]
    torch.cuda.synchronize()
    for b in range(nblk):
        single.run_device(d_in[b], out_s[b])
    b = 0
    for call in range(6):
        if call == 2:                                # a staged four-block and a staged two-block call in between
            octs.run_device_quad(d_in[b:b + 4], out_t[b:b + 4], staged=True)
            b += 4
            octs.run_device_pair(d_in[b], d_in[b + 1], out_t[b], out_t[b + 1], pipelined="staged")
            b += 2
        if call == 4:                                # single blocks (close the pipeline), then a JOINED eight-block call
            for _ in range(2):
                octs.run_device(d_in[b], out_t[b])
                b += 1
        octs.run_device_oct(d_in[b:b + 8], out_t[b:b + 8], staged=(call != 4))
        b += 8
        if call == 1:
            octs.join()
    assert b == nblk
    assert single.sync() == 0 and octs.sync() == 0
    assert single.blockcounter() == octs.blockcounter() == nblk
    for b in range(nblk):
        a, t = out_s[b].cpu().numpy(), out_t[b].cpu().numpy()
        if out_fmt in (8, 10):
            assert rel_rms(t.view(dt), a.view(dt)) < (2e-6 if rs == 4 else 1e-13), b
        else:
            dd = np.abs(decode_raw(a, out_fmt, n_out).ravel().astype(np.float64) - decode_raw(t, out_fmt, n_out).ravel())
            assert dd.max() <= 1, (b, dd.max())


def test_mac_profile_tells_the_launch_kinds_apart(pkg):
    """bfir_get_mac_profile: the partition-sum launches of a profiled run, by blocks per launch (8 / 4 / 2 / 1)."""
    import torch
    L, P, C, S = 2048, 3, 4, 24
    e = pkg.Brutefir(L, P, 8, C, pkg.FLOAT64_LE, pkg.FLOAT64_LE, 2000, False, n_streams=S, n_groups=1)
    assert e.set_coeff([decay_filter(c % 5, L * P) for c in range(C * S)], P) == 0
    d_in = [torch.rand(S * L * C, dtype=torch.float64, device="cuda") for _ in range(8)]
    d_out = [torch.empty(S * L * C, dtype=torch.float64, device="cuda") for _ in range(8)]
    torch.cuda.synchronize()
    for b in range(P):
        e.run_device(d_in[b], d_out[0])
    assert e.sync() == 0
    e.set_profiling(16)
    for nb in (1, 2, 4, 8):
        e.get_mac_profile(nb)
    e.run_device_oct(d_in, d_out)
    e.run_device_oct(d_in, d_out, staged=True)
    e.run_device_quad(d_in[:4], d_out[:4])
    e.run_device_pair(d_in[0], d_in[1], d_out[0], d_out[1])
    e.run_device(d_in[2], d_out[2])
    assert e.sync() == 0
    counts = {nb: e.get_mac_profile(nb) for nb in (8, 4, 2, 1)}
    assert [counts[nb][1] for nb in (8, 4, 2, 1)] == [2, 1, 1, 1], counts
    assert all(ms > 0 for ms, n in counts.values())
    assert e.get_mac_profile(8) == (0.0, 0)              # reset by the read above
    assert e.blockcounter() == P + 8 + 8 + 4 + 2 + 1


@pytest.mark.parametrize("rs,groups,P", [(8, 1, 5), (4, 1, 2), (4, 1, 7), (8, 3, 4)])
def test_host_quads_equal_single_blocks(pkg, rs, groups, P):
    """bfir_run_async_quad: four blocks of pinned host buffers per call through the stage pipeline (input copies, forward
    transforms, ONE four-block partition sum, inverse transforms, output copies on five streams over a ring of 12 staging
    slots), more calls in flight than the ring holds, mixed with two-block and one-block host calls and a device call;
    with several stream groups the call falls back to two pair calls. Same output as the synchronous bfir_run."""
    import torch
    L, C, S = 512, 2, 3
    fmt = pkg.FLOAT_LE if rs == 4 else pkg.FLOAT64_LE
    dt, tdt = (np.float32, torch.float32) if rs == 4 else (np.float64, torch.float64)
    h = [decay_filter(c, L * P) for c in range(C * S)]
    single = pkg.Brutefir(L, P, rs, C, fmt, fmt, 2000, False, n_streams=S, n_groups=1)
    quad = pkg.Brutefir(L, P, rs, C, fmt, fmt, 2000, False, n_streams=S, n_groups=groups)
    assert single.set_coeff(h, P) == 0 and quad.set_coeff(h, P) == 0
    nblk = P + 4 * 9 + 2 + 1 + 4
    x = white_noise(77, nblk * L, C * S).astype(dt)
    blocks = [np.ascontiguousarray(x[b * L:(b + 1) * L].reshape(L, S, C).transpose(1, 0, 2)).ravel() for b in range(nblk)]
    pin_in = [torch.from_numpy(b).pin_memory() for b in blocks]
    pin_out = [torch.zeros(S * L * C, dtype=tdt).pin_memory() for _ in range(nblk)]
    ref = []
    for b in range(nblk):
        rc, o = single.run(blocks[b].view(np.uint8))
        assert rc == 0
        ref.append(o.view(dt).copy())
    b, t = 0, None
    while b < P:                                     # fill the delay line block by block
        t = quad.run_async(pin_in[b].numpy(), pin_out[b].numpy())
        b += 1
    for call in range(9):
        if call == 4:                                # a two-block and a one-block host call between the four-block ones
            t = quad.run_async_pair(pin_in[b].numpy(), pin_in[b + 1].numpy(), pin_out[b].numpy(), pin_out[b + 1].numpy())
            b += 2
            t = quad.run_async(pin_in[b].numpy(), pin_out[b].numpy())
            b += 1
        t = quad.run_async_quad([p.numpy() for p in pin_in[b:b + 4]], [p.numpy() for p in pin_out[b:b + 4]])
        b += 4
        if call == 6:
            assert quad.wait(t) == 0
    d_in = [torch.from_numpy(blocks[k]).cuda() for k in range(b, b + 4)]
    d_out = [torch.zeros(S * L * C, dtype=tdt, device="cuda") for _ in range(4)]
    torch.cuda.synchronize()
    quad.run_device_quad(d_in, d_out, staged=True)   # a device call on the same pipeline
    assert quad.wait(t) == 0 and quad.sync() == 0
    for k in range(4):
        pin_out[b + k].copy_(d_out[k].cpu())
    b += 4
    assert b == nblk and quad.blockcounter() == nblk
    for k in range(nblk):
        assert rel_rms(pin_out[k].numpy(), ref[k]) < (2e-6 if rs == 4 else 1e-13), k


@pytest.mark.parametrize("rs,groups,P", [(4, 1, 6), (4, 3, 9), (8, 2, 5), (4, 2, 2)])
def test_block_quads_equal_single_blocks(pkg, rs, groups, P):
    """bfir_run_device_quad: four blocks per partition-sum launch (both precisions);
    same output as block by block up to the summation order, from the first block on (fall-back while filling)."""
    import torch
    L, C, S = 256, 2, 3
    fmt = pkg.FLOAT_LE if rs == 4 else pkg.FLOAT64_LE
    dt, tdt = (np.float32, torch.float32) if rs == 4 else (np.float64, torch.float64)
    h = [decay_filter(c, L * P - 3) for c in range(C * S)]
    single = pkg.Brutefir(L, P, rs, C, fmt, fmt, 2000, False, n_streams=S, n_groups=1)
    quad = pkg.Brutefir(L, P, rs, C, fmt, fmt, 2000, False, n_streams=S, n_groups=groups)
    assert single.set_coeff(h, P) == 0 and quad.set_coeff(h, P) == 0
    nblk = 28
    x = white_noise(61, nblk * L, C * S).astype(dt)
    d_in = [torch.from_numpy(np.ascontiguousarray(x[b * L:(b + 1) * L].reshape(L, S, C).transpose(1, 0, 2)).ravel()).cuda() for b in range(nblk)]
    n = S * L * C
    out_s = [torch.zeros(n, dtype=tdt, device="cuda") for _ in range(nblk)]
    out_q = [torch.zeros(n, dtype=tdt, device="cuda") for _ in range(nblk)]
    torch.cuda.synchronize()
    for b in range(nblk):
        single.run_device(d_in[b], out_s[b])
    for b in range(0, nblk, 4):
        quad.run_device_quad(d_in[b:b + 4], out_q[b:b + 4])
    assert single.sync() == 0 and quad.sync() == 0
    assert single.blockcounter() == quad.blockcounter() == nblk
    for b in range(nblk):
        assert rel_rms(out_q[b].cpu().numpy(), out_s[b].cpu().numpy()) < (2e-6 if rs == 4 else 1e-13), b


@pytest.mark.parametrize("rs", [4, 8])
def test_coefficient_set_routing(pkg, rs):
    """bfir_set_coeff_map (N4: the channels[] list of struct bfcoeff_t, global.h:71-78, which the reference always
    leaves at the identity): an engine whose filter channels share coefficient sets through a map produces the same
    bytes as an engine loaded with the explicitly duplicated filters -- one block per call (look-ahead head term),
    two and four blocks per call, staged calls and the synchronous host path; None restores the identity."""
    import torch
    L, P, C, S = 256, 4, 4, 2
    fmt = pkg.FLOAT_LE if rs == 4 else pkg.FLOAT64_LE
    dt, tdt = (np.float32, torch.float32) if rs == 4 else (np.float64, torch.float64)
    Ct = C * S
    h = [decay_filter(c, L * P - 5) for c in range(Ct)]
    mapping = [2, 2, 0, 1, 7, 5, 5, 4]
    mapped = pkg.Brutefir(L, P, rs, C, fmt, fmt, 2000, False, n_streams=S, n_groups=1)
    plain = pkg.Brutefir(L, P, rs, C, fmt, fmt, 2000, False, n_streams=S, n_groups=1)
    assert mapped.set_coeff(h, P) == 0 and plain.set_coeff([h[m] for m in mapping], P) == 0
    assert mapped.set_coeff_map(mapping) == 0
    with pytest.raises(pkg.BfirError):
        mapped.set_coeff_map(mapping[:-1])
    with pytest.raises(pkg.BfirError):
        mapped.set_coeff_map([Ct] * Ct)
    nblk = 30
    x = white_noise(71, nblk * L, Ct).astype(dt)
    blocks = [np.ascontiguousarray(x[b * L:(b + 1) * L].reshape(L, S, C).transpose(1, 0, 2)).ravel() for b in range(nblk)]
    d_in = [torch.from_numpy(b).cuda() for b in blocks]
    n = S * L * C
    out = {k: [torch.zeros(n, dtype=tdt, device="cuda") for _ in range(nblk)] for k in ("m", "p")}
    torch.cuda.synchronize()
    for name, e in (("m", mapped), ("p", plain)):
        o = out[name]
        b = 0
        for _ in range(6):
            e.run_device(d_in[b], o[b]); b += 1
        rc, y = e.run(blocks[b].view(np.uint8))                      # synchronous host path (look-ahead head term)
        assert rc == 0
        o[b].copy_(torch.from_numpy(y.view(dt))); b += 1
        rc, y = e.run(blocks[b].view(np.uint8))
        assert rc == 0
        o[b].copy_(torch.from_numpy(y.view(dt))); b += 1
        e.run_device_pair(d_in[b], d_in[b + 1], o[b], o[b + 1]); b += 2
        e.run_device_quad(d_in[b:b + 4], o[b:b + 4]); b += 4
        e.run_device_quad(d_in[b:b + 4], o[b:b + 4], staged=True); b += 4
        e.run_device_pair(d_in[b], d_in[b + 1], o[b], o[b + 1], pipelined="staged"); b += 2
        e.join()
        while b < nblk:
            e.run_device(d_in[b], o[b]); b += 1
        assert e.sync() == 0
    for b in range(nblk):
        assert torch.equal(out["m"][b], out["p"][b]), b
    # back to the identity: now the two engines differ (different filters per channel)
    assert mapped.set_coeff_map(None) == 0
    ym, yp = torch.zeros(n, dtype=tdt, device="cuda"), torch.zeros(n, dtype=tdt, device="cuda")
    mapped.run_device(d_in[0], ym)
    plain.run_device(d_in[0], yp)
    assert mapped.sync() == 0 and plain.sync() == 0
    assert not torch.equal(ym, yp)
