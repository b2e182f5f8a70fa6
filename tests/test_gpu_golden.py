"""GPU parity against the committed golden vectors (tests/golden/golden_v1.npz, frozen from the
reference's own sources by tests/golden/make_golden.py) -- reads no oracle and no /root/reference."""
import os
import zlib

import numpy as np
import pytest

from conftest import rel_rms, parity

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz")
TOL = {4: 1e-5, 8: 1e-12}
TAGS = {4: "f32", 8: "f64"}


@pytest.fixture(scope="module")
def G():
    return np.load(GOLDEN)


@pytest.mark.parametrize("rs", [4, 8])
def test_convolver_entry_points(pkg, G, rs):
    L, tag = 64, TAGS[rs]
    k = "conv/%s/" % tag
    g = pkg.FftwConvolver(L, rs, 2, 2000)
    up = lambda name: g.cbuf(G[k + name])
    hcs = []
    for name in ("x", "x2", "x3"):
        b = g.cbuf()
        g.convolver_time2freq(up(name), b)
        hcs.append(b)
    assert rel_rms(g.get(hcs[0]), G[k + "time2freq"]) < TOL[rs]
    t = g.cbuf()
    g.convolver_freq2time(g.cbuf(G[k + "time2freq"]), t)
    assert rel_rms(g.get(t), G[k + "freq2time"]) < TOL[rs]
    # element-wise entry points on the golden inputs: bit-exact
    ghc = [g.cbuf(G[k + "time2freq"])]
    out = g.cbuf()
    g.convolver_mixnscale(ghc, out, [0.5], 1)
    assert np.array_equal(g.get(out), G[k + "mix_in_1"])
    o1 = g.cbuf()
    g.convolver_mixnscale(ghc, o1, [1.0], 1)
    g.convolver_mixnscale([o1], out, [3.0], 3)
    assert np.array_equal(g.get(out), G[k + "mix_out_1"])
    c = g.cbuf()
    assert g.convolver_coeffs2cbuf(G[k + "h"], len(G[k + "h"]), 0.75, c) == 0
    assert rel_rms(g.get(c), G[k + "coeffs2cbuf"]) < TOL[rs]
    gc = g.cbuf(G[k + "coeffs2cbuf"])
    g.convolver_convolve(o1, gc, out)
    assert np.array_equal(g.get(out), G[k + "convolve"])
    g.convolver_dirac_convolve(ghc[0], out)
    assert np.array_equal(g.get(out), G[k + "dirac"])
    if rs == 4:
        bi, bx, bb = g.cbuf(G[k + "convolve"]), g.cbuf(G[k + "xfade_old"]), g.cbuf()
        g.convolver_crossfade_inplace(bi, bx, bb)
        parity("golden/crossfade_inplace/f32", g.get(bi), G[k + "crossfade"], TOL[rs])
    buf = g.cbuf(n_cbufs=1.5)
    for i, name in enumerate(("x", "x2", "x3")):
        g.convolver_convolve_eval(hcs[i], buf, out)
        parity("golden/convolve_eval/%s/%d" % (tag, i), g.get(out), G[k + "convolve_eval"][i], TOL[rs])


@pytest.mark.parametrize("rs", [4, 8])
def test_td_convolver(pkg, G, rs):
    k = "conv/%s/" % TAGS[rs]
    g = pkg.FftwConvolver(64, rs, 2, 2000)
    t = g.convolver_td_new(G[k + "td_h"])
    assert rel_rms(g.convolver_td_coeffs(t), G[k + "td_coeffs"]) < TOL[rs]
    d = g.rawbuf(G[k + "td_x"])
    g.convolver_td_convolve(t, d)
    parity("golden/td_convolve/%s" % TAGS[rs], d.download(g.dtype), G[k + "td_convolve"], TOL[rs])
    g.convolver_td_free(t)


@pytest.mark.parametrize("rs", [4, 8])
def test_codecs_and_dither(pkg, G, rs):
    L, C, tag = 64, 3, TAGS[rs]
    g = pkg.FftwConvolver(L, rs, 2, 2000)
    for fmt in range(1, 12):
        nbytes = pkg.FORMAT_BYTES[fmt]
        draw = g.rawbuf(G["codec/%s/raw_in_%d" % (tag, fmt)])
        for ch in range(C):
            cb, nb = g.cbuf(), g.cbuf()
            g.convolver_raw2cbuf(draw, cb, nb, fmt, ch * nbytes, C)
            assert np.array_equal(g.get(nb)[:L], G["codec/%s/raw2real_%d" % (tag, fmt)][ch])
        ov = pkg.Overflow()
        ov.max = 1.0 if fmt >= 8 else float(2 ** (8 * nbytes - 1)) - 1
        dout = g.rawbuf(nbytes=L * nbytes)
        g.convolver_cbuf2raw(g.cbuf(G["codec/%s/real_in_%d" % (tag, fmt)]), dout, fmt, 0, 1, False, 0, ov)
        assert np.array_equal(dout.download(np.uint8), G["codec/%s/real2raw_%d" % (tag, fmt)])
        assert np.array_equal(np.array(ov.as_tuple()), G["codec/%s/real2raw_overflow_%d" % (tag, fmt)])
    tab = g.dither_table()
    assert np.array_equal(tab[:4096], G["dither/%s/table_head" % tag])
    assert [len(tab), zlib.crc32(tab.tobytes())] == list(G["dither/%s/table_size_crc" % tag])
    assert np.array_equal(g.dither_map()[:511], G["dither/%s/map" % tag])
    ov = pkg.Overflow()
    ov.max = 32767.0
    dout = g.rawbuf(nbytes=L * 2)
    for blk in range(3):
        g.convolver_cbuf2raw(g.cbuf(G["dither/%s/real_in" % tag][blk]), dout, pkg.S16_LE, 0, 1, True, 1, ov)
        assert np.array_equal(dout.download(np.uint8), G["dither/%s/s16_out" % tag][blk])
        assert g.dither_ptr(1) == G["dither/%s/ptrs" % tag][blk]
    assert np.array_equal(np.array(ov.as_tuple()), G["dither/%s/overflow" % tag])


@pytest.mark.parametrize("rs", [4, 8])
def test_engine(pkg, G, rs):
    L, P, C, tag = 64, 3, 2, TAGS[rs]
    fmt = pkg.FLOAT_LE if rs == 4 else pkg.FLOAT64_LE
    dt = np.float32 if rs == 4 else np.float64
    h, xin = list(G["engine/%s/h" % tag]), G["engine/%s/x" % tag]
    for name, out_fmt, dith in (("float", fmt, False), ("s16", pkg.S16_LE, False), ("s16_dither", pkg.S16_LE, True)):
        e = pkg.Brutefir(L, P, rs, C, fmt, out_fmt, 2000, dith)
        assert e.set_coeff(h, P, 0.9) == 0
        for b in range(8):
            raw = np.ascontiguousarray(xin[b * L:(b + 1) * L].astype(dt)).view(np.uint8).ravel()
            rc, out = e.run(raw)
            assert rc == 0
            ref = G["engine/%s/out_%s" % (tag, name)][b]
            if name == "float":
                assert rel_rms(out.view(dt), ref.view(dt)) < TOL[rs]
            else:
                d = np.abs(out.view("<i2").astype(int) - ref.view("<i2").astype(int))
                assert d.max() <= (1 if not dith else 3)
        if dith:
            assert [e.dither_ptr(c) for c in range(C)] == list(G["engine/%s/dither_ptrs" % tag])
