"""CPU tests that pin the oracle (TEST INFRASTRUCTURE): the C++ restatement (oracle/bfir_oracle.cpp,
"port") and -- where /root/reference or a prebuilt oracle/_ref exists -- the unmodified reference
sources ("ref") are checked against
  * the golden vectors frozen from the reference build (tests/golden/golden_v1.npz),
  * the independent numpy float64 restatement (oracle/oracle_np.py), and
  * mathematical properties (run() == direct linear convolution, HC2R(R2HC(x)) = N x, ...),
so that GPU parity against either oracle means parity with the reference."""
import os
import zlib

import numpy as np
import pytest

from conftest import oracle_kinds, rel_rms, decay_filter, white_noise, encode_raw, decode_raw

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz")
TOL = {4: 2e-6, 8: 4e-15}
TAGS = {4: "f32", 8: "f64"}


@pytest.fixture(scope="module")
def G():
    return np.load(GOLDEN)


@pytest.mark.parametrize("kind", oracle_kinds())
@pytest.mark.parametrize("rs", [4, 8])
def test_convolver_entry_points_match_golden(oracle, G, kind, rs):
    L, tag = 64, TAGS[rs]
    k = "conv/%s/" % tag
    cv = oracle.Convolver(L, rs, kind)
    x, x2, x3 = G[k + "x"], G[k + "x2"], G[k + "x3"]
    exact = (lambda a, b: np.array_equal(a, b))
    hc, hc2, hc3 = cv.time2freq(x), cv.time2freq(x2), cv.time2freq(x3)
    assert exact(hc, G[k + "time2freq"])               # same FFT provider in both oracle builds
    assert exact(cv.freq2time(hc.copy()), G[k + "freq2time"])
    assert exact(cv.mixnscale([hc], [0.5], 1), G[k + "mix_in_1"])
    sc = list(G[k + "mix_scales"])
    assert exact(cv.mixnscale([hc, hc2, hc3], sc, 1), G[k + "mix_in_3"])
    o1, o2, o3 = (cv.mixnscale([h], [1.0], 1) for h in (hc, hc2, hc3))
    assert exact(cv.mixnscale([o1], [3.0], 3), G[k + "mix_out_1"])
    assert exact(cv.mixnscale([o1, o2, o3], sc, 3), G[k + "mix_out_3"])
    c = cv.coeffs2cbuf(G[k + "h"], 0.75)
    assert exact(c, G[k + "coeffs2cbuf"])
    assert exact(cv.runtime_coeffs2cbuf(np.concatenate([G[k + "h"], np.zeros(5, dtype=cv.dtype)])), G[k + "runtime_coeffs2cbuf"])
    conv = cv.convolve(o1, c)
    assert exact(conv, G[k + "convolve"])
    assert exact(cv.convolve_add(o2, c, conv.copy()), G[k + "convolve_add"])
    assert exact(cv.convolve_inplace(o3.copy(), c), G[k + "convolve_inplace"])
    assert exact(cv.dirac_convolve(hc), G[k + "dirac"])
    if rs == 4:
        assert exact(cv.crossfade_inplace(conv.copy(), G[k + "xfade_old"].copy(), cv.cbuf()), G[k + "crossfade"])
    buf = np.zeros(3 * L, dtype=cv.dtype)
    for i, h_ in enumerate((hc, hc2, hc3)):
        assert exact(cv.convolve_eval(h_, buf), G[k + "convolve_eval"][i])
    assert exact(buf, G[k + "convolve_eval_buffer"])
    assert exact(cv.preprocess_coeff(G[k + "preprocess_coeff_in"], 4, 0.5), G[k + "preprocess_coeff"])
    bl, tdc = cv.td_new(G[k + "td_h"])                  # td_conv_t, fftw_convolver.cpp:698-777
    assert bl == 32 == len(G[k + "td_x"]) // 2
    assert exact(cv.td_coeffs(tdc, bl), G[k + "td_coeffs"])
    assert exact(cv.td_convolve(tdc, G[k + "td_x"].copy()), G[k + "td_convolve"])
    cv.td_free(tdc)


@pytest.mark.parametrize("kind", oracle_kinds())
@pytest.mark.parametrize("rs", [4, 8])
def test_codecs_match_golden(oracle, G, kind, rs):
    L, C, tag = 64, 3, TAGS[rs]
    cv = oracle.Convolver(L, rs, kind, n_channels=2, sample_rate=2000)
    for fmt in range(1, 12):
        raw = G["codec/%s/raw_in_%d" % (tag, fmt)]
        for ch in range(C):
            cb, nb = cv.cbuf(), cv.cbuf()
            cv.raw2cbuf(raw, cb, nb, fmt, ch, C)
            assert np.array_equal(nb[:L], G["codec/%s/raw2real_%d" % (tag, fmt)][ch])
            assert np.array_equal(cb[L:], nb[:L]) and np.all(cb[:L] == 0)
        nbytes = oracle.FORMAT_BYTES[fmt]
        ov = oracle.Overflow()
        ov.max = 1.0 if fmt >= 8 else float(2 ** (8 * nbytes - 1)) - 1
        out = np.zeros(L * nbytes, dtype=np.uint8)
        cv.cbuf2raw(G["codec/%s/real_in_%d" % (tag, fmt)], out, fmt, 0, 1, False, 0, ov)
        assert np.array_equal(out, G["codec/%s/real2raw_%d" % (tag, fmt)])
        assert np.array_equal(np.array(ov.as_tuple()), G["codec/%s/real2raw_overflow_%d" % (tag, fmt)])


def test_quantiser_truth_table(oracle):
    """mid-tread requantiser of dither.cpp:215-274: trunc-then-decrement for negatives, clip at the limits"""
    cv = oracle.Convolver(16, 8, "port")
    y = np.zeros(32)
    y[:11] = [0.0, 0.49, 0.5, -0.49, -0.5, -0.51, 3.0, -3.0, 32767.0, 32767.6, -32768.6]
    ov = oracle.Overflow()
    ov.max = 32767.0
    out = np.zeros(32, dtype=np.uint8)
    cv.cbuf2raw(y, out, oracle.S16_LE, 0, 1, False, 0, ov)
    got = out.view("<i2")[:11]
    # -0.49 + 0.5 = +0.01 -> 0;  -0.5 + 0.5 = 0 -> 0;  -0.51 -> trunc(-0.01) - 1 = -1;
    # -3.0 -> trunc(-2.5) - 1 = -3;  32767.6 + 0.5 > rmax -> clip;  -32768.6 + 0.5 <= rmin -> clip
    assert list(got) == [0, 0, 1, 0, 0, -1, 3, -3, 32767, 32767, -32768]
    # 32767.0 + 0.5 > rmax as well: a full-scale sample is clipped to itself but COUNTED (dither.cpp:252-256)
    assert ov.n_overflows == 3


@pytest.mark.parametrize("kind", oracle_kinds())
@pytest.mark.parametrize("rs", [4, 8])
def test_dither_matches_golden(oracle, G, kind, rs):
    L, tag = 64, TAGS[rs]
    cv = oracle.Convolver(L, rs, kind, n_channels=2, sample_rate=2000)
    tab = cv.dither_table()
    assert np.array_equal(tab[:4096], G["dither/%s/table_head" % tag])
    assert [len(tab), zlib.crc32(tab.tobytes())] == list(G["dither/%s/table_size_crc" % tag])
    assert np.array_equal(cv.dither_map(), G["dither/%s/map" % tag])
    ov = oracle.Overflow()
    ov.max = 32767.0
    for blk in range(3):
        out = np.zeros(L * 2, dtype=np.uint8)
        cv.cbuf2raw(G["dither/%s/real_in" % tag][blk], out, oracle.S16_LE, 0, 1, True, 1, ov)
        assert np.array_equal(out, G["dither/%s/s16_out" % tag][blk])
        assert cv.dither_ptr(1) == G["dither/%s/ptrs" % tag][blk]
    assert np.array_equal(np.array(ov.as_tuple()), G["dither/%s/overflow" % tag])


def test_tausworthe_known_answer(oracle):
    """GSL `taus` generator (dither.cpp:411-449): state after seeding with 1 and six warm-up draws; the
    table bytes are the low bytes of the following draws. Restated here independently in pure Python."""
    def step(s, a, b, c, d):
        return (((s & c) << d) & 0xFFFFFFFF) ^ ((((s << a) & 0xFFFFFFFF) ^ s) >> b)
    s = [69069 & 0xFFFFFFFF, 0, 0]
    s[1] = (69069 * s[0]) & 0xFFFFFFFF
    s[2] = (69069 * s[1]) & 0xFFFFFFFF

    def draw():
        s[0] = step(s[0], 13, 19, 4294967294, 12)
        s[1] = step(s[1], 2, 25, 4294967288, 4)
        s[2] = step(s[2], 3, 11, 4294967280, 17)
        return s[0] ^ s[1] ^ s[2]
    for _ in range(6):
        draw()
    want = np.array([draw() & 0xFF for _ in range(256)], dtype=np.uint8).view(np.int8)
    for kind in oracle_kinds():
        tab = oracle.Convolver(64, 4, kind, n_channels=1, sample_rate=1000).dither_table()
        assert np.array_equal(tab[:256], want)
        assert len(tab) == 1 * 10 * 1000 + 1          # n_channels * RANDTAB_SPACING * rate + 1


@pytest.mark.parametrize("kind", oracle_kinds())
@pytest.mark.parametrize("rs", [4, 8])
def test_engine_matches_golden(oracle, G, kind, rs):
    L, P, C, tag = 64, 3, 2, TAGS[rs]
    fmt = oracle.FLOAT_LE if rs == 4 else oracle.FLOAT64_LE
    h, xin = list(G["engine/%s/h" % tag]), G["engine/%s/x" % tag]
    dt = np.float32 if rs == 4 else np.float64
    for name, out_fmt, dith in (("float", fmt, False), ("s16", oracle.S16_LE, False), ("s16_dither", oracle.S16_LE, True)):
        e = oracle.Engine(L, P, rs, C, fmt, out_fmt, 2000, dith, kind=kind)
        assert e.set_coeff(h, P, 0.9) == 0
        for b in range(8):
            raw = np.ascontiguousarray(xin[b * L:(b + 1) * L].astype(dt)).view(np.uint8).ravel()
            rc, out = e.run(raw)
            assert rc == 0
            assert np.array_equal(out, G["engine/%s/out_%s" % (tag, name)][b]), (name, b)
        ov = np.array([e.overflow(c).as_tuple() for c in range(C)], dtype=np.float64)
        assert np.array_equal(ov, G["engine/%s/overflow_%s" % (tag, name)])
        if dith:
            assert [e.dither_ptr(c) for c in range(C)] == list(G["engine/%s/dither_ptrs" % tag])


@pytest.mark.parametrize("kind", oracle_kinds())
@pytest.mark.parametrize("rs", [4, 8])
def test_equalizer_matches_golden(oracle, G, kind, rs):
    out = oracle.equalizer_render(64, 16, rs, 48000, G["eq/bands"], G["eq/mag_db"], G["eq/phase"], kind=kind)
    ref = G["eq/%s/render" % TAGS[rs]]
    assert rel_rms(out, ref) < (1e-5 if rs == 4 else 1e-12)   # libm sincos vs the frozen values


@pytest.mark.parametrize("kind", oracle_kinds())
@pytest.mark.parametrize("rs", [4, 8])
def test_against_numpy_restatement(oracle, kind, rs):
    from oracle import oracle_np as onp
    L = 256
    cv = oracle.Convolver(L, rs, kind)
    rng = np.random.default_rng(rs)
    x = rng.uniform(-1, 1, 2 * L).astype(cv.dtype)
    hc = cv.time2freq(x)
    assert rel_rms(hc, onp.r2hc(x)) < TOL[rs]
    assert rel_rms(cv.freq2time(hc.copy()), 2 * L * x.astype(np.float64)) < TOL[rs]      # HC2R(R2HC(x)) = N x
    o = cv.mixnscale([hc], [0.25], 1)
    assert np.allclose(o, onp.hc_to_ord(hc, 0.25), rtol=0, atol=0)
    assert np.array_equal(cv.mixnscale([o], [4.0], 3), hc)                                # OUTPUT o INPUT = id
    h = rng.standard_normal(L - 9).astype(cv.dtype)
    c = cv.coeffs2cbuf(h)
    assert rel_rms(c, onp.coeffs2cbuf(h, L)) < TOL[rs]
    assert rel_rms(cv.convolve(o, c), onp.convolve_ord(o, c)) < TOL[rs]
    # dirac_convolve == convolution with a unit impulse (on HC buffers)
    one = cv.coeffs2cbuf(np.array([1.0]))
    via = cv.mixnscale([cv.convolve(cv.mixnscale([hc], [1.0], 1), one)], [1.0], 3)
    assert rel_rms(cv.dirac_convolve(hc), via) < TOL[rs] * 4


@pytest.mark.parametrize("kind", oracle_kinds())
@pytest.mark.parametrize("rs,L,P,C", [(4, 32, 5, 2), (8, 32, 5, 2), (8, 128, 1, 1), (4, 256, 9, 3)])
def test_run_equals_direct_linear_convolution(oracle, kind, rs, L, P, C):
    from oracle import oracle_np as onp
    fmt = oracle.FLOAT_LE if rs == 4 else oracle.FLOAT64_LE
    dt = np.float32 if rs == 4 else np.float64
    e = oracle.Engine(L, P, rs, C, fmt, fmt, 44100, False, kind=kind)
    h = [decay_filter(c, L * P - 7).astype(dt) for c in range(C)]
    assert e.set_coeff(h, P) == 0
    nb = 3 * P + 1
    x = white_noise(1, nb * L, C).astype(dt)
    ys = []
    for b in range(nb):
        rc, out = e.run(np.ascontiguousarray(x[b * L:(b + 1) * L]).view(np.uint8).ravel())
        assert rc == 0
        ys.append(out.view(dt).reshape(L, C))
    y = np.concatenate(ys)
    np_e = onp.EngineNP(L, P, C)
    np_e.set_coeff(h, P)
    y_np = np.concatenate([np_e.run(x[b * L:(b + 1) * L].T.astype(np.float64)).T for b in range(nb)])
    for c in range(C):
        assert rel_rms(y[:, c], onp.direct_convolution(x[:, c], h[c], nb * L)) < (1e-6 if rs == 4 else 1e-13)
        assert rel_rms(y_np[:, c], onp.direct_convolution(x[:, c], h[c], nb * L)) < 1e-13


def test_ref_and_port_agree_on_a_long_run(oracle):
    if "ref" not in oracle_kinds():
        pytest.skip("oracle/_ref not available")
    L, P, C = 512, 6, 8
    a = oracle.Engine(L, P, 8, C, oracle.S24_LE, oracle.S32_BE, 48000, True, kind="ref")
    b = oracle.Engine(L, P, 8, C, oracle.S24_LE, oracle.S32_BE, 48000, True, kind="port")
    h = [decay_filter(c, L * P) for c in range(C)]
    assert a.set_coeff(h, P, 1.5) == 0 and b.set_coeff(h, P, 1.5) == 0
    x = white_noise(4, 20 * L, C)
    for blk in range(20):
        raw = encode_raw(x[blk * L:(blk + 1) * L], oracle.S24_LE)
        ra, ya = a.run(raw)
        rb, yb = b.run(raw)
        assert ra == rb == 0 and np.array_equal(ya, yb)
    for c in range(C):
        assert a.overflow(c).as_tuple() == b.overflow(c).as_tuple()
        assert a.dither_ptr(c) == b.dither_ptr(c)


def test_nonfinite_and_edge_cases(oracle):
    for kind in oracle_kinds():
        cv = oracle.Convolver(64, 4, kind)
        bad = np.ones(64, dtype=np.float32)
        bad[3] = np.inf
        assert cv.coeffs2cbuf(bad) is None                   # fftw_convolver.cpp:493-497
        assert np.all(cv.coeffs2cbuf(np.zeros(1, dtype=np.float32), n=0) == 0)
        e = oracle.Engine(64, 2, 4, 2, oracle.FLOAT_LE, oracle.FLOAT_LE, 44100, False, kind=kind)
        assert not e.is_initialized()
        hb = [np.ones(128), np.ones(128)]
        hb[1][7] = np.nan
        if kind == "ref":
            # reference quirk: coeff::preprocess_coeff only logs a failed block (coeff.cpp:345-348) and
            # returns a table holding a NULL partition, so brutefir::set_coeff reports success
            # (brutefir.cpp:207-227) and the next run() would dereference NULL. The build (and the port)
            # return the documented -2 instead; see DESIGN.md "reference quirks".
            assert e.set_coeff(hb, 2) == 0
            e = oracle.Engine(64, 2, 4, 2, oracle.FLOAT_LE, oracle.FLOAT_LE, 44100, False, kind=kind)
        else:
            assert e.set_coeff(hb, 2) == -2 and not e.is_initialized()
        assert e.set_coeff([np.ones(128), np.ones(128)], 2) == 0 and e.is_initialized()
        x = np.zeros((64, 2), dtype=np.float32)
        x[5, 0] = np.nan
        assert e.run(x.view(np.uint8).ravel())[0] == -1 and e.blockcounter() == 0


@pytest.mark.parametrize("kind", oracle_kinds())
@pytest.mark.parametrize("rs", [4, 8])
def test_td_convolver_is_a_circular_convolution(oracle, kind, rs):
    """td_conv_t (fftw_convolver.cpp:698-777): block (x) [0_blocklen | h | 0], circular over 2 * blocklen."""
    cv = oracle.Convolver(64, rs, kind)
    assert [cv.td_block_length(n) for n in (0, 1, 2, 3, 4, 5, 31, 32, 33)] == [-1, -1, 2, 4, 4, 8, 32, 32, 64]
    rng = np.random.default_rng(5)
    for n in (2, 7, 31, 64, 100):
        h = rng.standard_normal(n).astype(cv.dtype)
        bl, tdc = cv.td_new(h)
        x = rng.uniform(-1, 1, 2 * bl).astype(cv.dtype)
        hp = np.zeros(2 * bl)
        hp[bl:bl + n] = h
        want = np.real(np.fft.ifft(np.fft.fft(x.astype(np.float64)) * np.fft.fft(hp)))
        assert rel_rms(cv.td_convolve(tdc, x.copy()), want) < (2e-6 if rs == 4 else 1e-14)
        cv.td_free(tdc)


def test_mkl_fft_provider_for_the_cpu_baseline(oracle):
    """The baseline-timing build of the reference (oracle/_ref/libbfir_ref_mkl.so: the same unmodified sources, MKL
    DFTI out of libtorch_cpu.so behind the FFTW calls, SURVEY.md 8c provider 2) computes the same thing as the
    build the parity tests use: transforms against numpy, a whole run() against the first provider."""
    if not (oracle.available("ref_mkl") and oracle.best_timing_kind() == "ref_mkl"):
        pytest.skip("oracle/_ref/libbfir_ref_mkl.so or libtorch_cpu.so not present")
    from oracle import oracle_np as onp
    assert b"MKL" in oracle.lib("ref_mkl").fft_provider()
    for rs in (4, 8):
        for L in (16, 512, 8192):
            cv = oracle.Convolver(L, rs, kind="ref_mkl")
            x = np.random.default_rng(L).uniform(-1, 1, 2 * L).astype(cv.dtype)
            hc = cv.time2freq(x.copy())
            assert rel_rms(hc, onp.r2hc(x)) < TOL[rs]
            assert rel_rms(cv.freq2time(hc.copy()), 2 * L * x.astype(np.float64)) < TOL[rs]
            buf = x.copy()                       # in place, as the reference calls it (fftw_convolver.cpp:193-200)
            cv.time2freq(buf, buf)
            assert np.array_equal(buf, hc)
    L, P, C = 256, 6, 3
    a = oracle.Engine(L, P, 8, C, oracle.FLOAT64_LE, oracle.FLOAT64_LE, 48000, False, kind="ref")
    b = oracle.Engine(L, P, 8, C, oracle.FLOAT64_LE, oracle.FLOAT64_LE, 48000, False, kind="ref_mkl")
    h = [decay_filter(c, L * P) for c in range(C)]
    assert a.set_coeff(h, P) == 0 and b.set_coeff(h, P) == 0
    x = white_noise(4, 10 * L, C)
    for blk in range(10):
        raw = np.ascontiguousarray(x[blk * L:(blk + 1) * L]).view(np.uint8).ravel()
        ya, yb = a.run(raw)[1].view(np.float64), b.run(raw)[1].view(np.float64)
        assert rel_rms(yb, ya) < 1e-13


@pytest.mark.parametrize("rs", [4, 8])
def test_reference_offline_drivers_in_the_oracle(oracle, rs):
    """brutefir/preprocessor.cpp compiled into oracle/_ref (in-memory sound files instead of libsndfile) behaves as
    its source reads: the cascade keeps only the FIRST block of each intermediate result as the next coefficient set
    (set_coeff is handed filter_length as the coefficient length, preprocessor.cpp:176), scaled by the file's scale;
    the attenuation probe is -20 log10 of the peak of noise * first block of the response."""
    if not oracle.available("ref"):
        pytest.skip("oracle/_ref not built")
    L, C = 64, 2
    rng = np.random.default_rng(3)
    dt = np.float32 if rs == 4 else np.float64
    irs = [(rng.standard_normal((f, C)) * np.exp(-np.arange(f) / 50.0)[:, None] * 0.3).astype(dt) for f in (150, 200, 90)]
    scales = [1.0, 0.5, 2.0]
    y = oracle.ref_convolve_impulses(irs, scales, L, rs)
    assert y.shape == (200, C)
    T = L * ((200 + L - 1) // L)
    want = np.zeros((T, C))
    for c in range(C):
        h = np.zeros(L)
        h[0] = 1.0
        carry = np.zeros(0)
        for k, ir in enumerate(irs):
            x = np.zeros(T)
            x[:ir.shape[0]] = ir[:, c]
            stream = np.concatenate([carry, x])                   # the engine is never reset: the delay line carries on
            out = np.convolve(stream, h)[len(carry):len(carry) + T]
            carry = stream
            h = out[:L] * scales[k]
        want[:, c] = out
    assert rel_rms(y, want[:200]) < (2e-6 if rs == 4 else 1e-13)
    resp = irs[1] * 10
    noise = rng.uniform(-1, 1, T * C).astype(dt)
    att = oracle.ref_calculate_attenuation(resp, L, rs, noise)
    nz = noise.reshape(T, C).astype(np.float64)
    peak = max(np.abs(np.convolve(nz[:, c], resp[:L, c].astype(np.float64))[:T]).max() for c in range(C))
    assert peak > 1 and abs(att + 20 * np.log10(peak)) < (1e-4 if rs == 4 else 1e-9)
