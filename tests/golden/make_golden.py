"""Generates tests/golden/golden_v1.npz from the UNMODIFIED reference sources (oracle/_ref, i.e.
/root/reference/brutefir compiled in place through oracle/ref_shim, FFT provider oracle/fft_r2r).

The reference ships no golden vectors or known-answer tests; these vectors are
outputs of the reference's own code run in the build container, frozen so that the GPU box (where
/root/reference does not exist) can check both the CUDA path and the oracle restatement against them.

    python tests/golden/make_golden.py          # needs /root/reference (builds oracle/_ref on demand)

Entries are keyed "<group>/<precision>/<name>"; inputs are stored next to outputs.
"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
FORMATS = list(range(1, 12))


def encode_raw(x, fmt):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import encode_raw as enc
    return enc(x, fmt)


def main():
    if not oracle.available("ref"):
        oracle.build("ref")
    assert oracle.available("ref"), "oracle/_ref could not be built (is /root/reference present?)"
    G = {}
    rng = np.random.default_rng(20261018)
    L = 64
    for rs, tag in ((4, "f32"), (8, "f64")):
        cv = oracle.Convolver(L, rs, "ref", n_channels=2, sample_rate=2000)
        dt = cv.dtype
        k = "conv/%s/" % tag
        x = rng.uniform(-1, 1, 2 * L).astype(dt)
        x2 = rng.uniform(-1, 1, 2 * L).astype(dt)
        x3 = rng.uniform(-1, 1, 2 * L).astype(dt)
        G[k + "x"], G[k + "x2"], G[k + "x3"] = x, x2, x3
        hc = cv.time2freq(x)
        G[k + "time2freq"] = hc
        G[k + "freq2time"] = cv.freq2time(hc.copy())
        G[k + "mix_in_1"] = cv.mixnscale([hc], [0.5], oracle.MIX_INPUT)
        hc2, hc3 = cv.time2freq(x2), cv.time2freq(x3)
        scales = [0.5, -1.25, 2.0]
        G[k + "mix_scales"] = np.array(scales)
        G[k + "mix_in_3"] = cv.mixnscale([hc, hc2, hc3], scales, oracle.MIX_INPUT)
        o1, o2, o3 = (cv.mixnscale([h], [1.0], oracle.MIX_INPUT) for h in (hc, hc2, hc3))
        G[k + "mix_out_1"] = cv.mixnscale([o1], [3.0], oracle.MIX_OUTPUT)
        G[k + "mix_out_3"] = cv.mixnscale([o1, o2, o3], scales, oracle.MIX_OUTPUT)
        h = (rng.standard_normal(L - 5) * np.exp(-np.arange(L - 5) / 16.0)).astype(dt)
        G[k + "h"] = h
        c = cv.coeffs2cbuf(h, 0.75)
        G[k + "coeffs2cbuf"] = c
        G[k + "runtime_coeffs2cbuf"] = cv.runtime_coeffs2cbuf(np.concatenate([h, np.zeros(5, dtype=dt)]))
        G[k + "convolve"] = cv.convolve(o1, c)
        G[k + "convolve_add"] = cv.convolve_add(o2, c, G[k + "convolve"].copy())
        G[k + "convolve_inplace"] = cv.convolve_inplace(o3.copy(), c)
        G[k + "dirac"] = cv.dirac_convolve(hc)
        if rs == 4:  # the reference's double branch reads memory it never wrote (fftw_convolver.cpp:306-315)
            old = cv.convolve(o1, cv.coeffs2cbuf(rng.standard_normal(L).astype(dt)))
            G[k + "xfade_old"] = old
            G[k + "crossfade"] = cv.crossfade_inplace(G[k + "convolve"].copy(), old.copy(), cv.cbuf())
        buf = np.zeros(3 * L, dtype=dt)
        ev = []
        for h_ in (hc, hc2, hc3):
            ev.append(cv.convolve_eval(h_, buf).copy())
        G[k + "convolve_eval"] = np.stack(ev)
        G[k + "convolve_eval_buffer"] = buf.copy()
        pc_in = rng.standard_normal(3 * L - 7).astype(dt)
        G[k + "preprocess_coeff_in"] = pc_in
        G[k + "preprocess_coeff"] = cv.preprocess_coeff(pc_in, 4, 0.5)

        # td_conv_t: small one-shot convolver on the plain HC layout (fftw_convolver.cpp:698-777)
        td_rng = np.random.default_rng(31 + rs)  # own generator: the vectors above and below keep their draws
        td_h = (td_rng.standard_normal(31) * np.exp(-np.arange(31) / 8.0)).astype(dt)
        bl, tdc = cv.td_new(td_h)
        td_x = td_rng.uniform(-1, 1, 2 * bl).astype(dt)
        G[k + "td_h"], G[k + "td_x"] = td_h, td_x
        G[k + "td_coeffs"] = cv.td_coeffs(tdc, bl)
        G[k + "td_convolve"] = cv.td_convolve(tdc, td_x.copy())
        cv.td_free(tdc)

        # codecs
        C = 3
        xr = rng.uniform(-1, 1, (L, C))
        for fmt in FORMATS:
            raw = encode_raw(xr, fmt)
            G["codec/%s/raw_in_%d" % (tag, fmt)] = raw
            outs = []
            for ch in range(C):
                cb, nb = cv.cbuf(), cv.cbuf()
                cv.raw2cbuf(raw, cb, nb, fmt, ch, C)
                outs.append(nb[:L].copy())
            G["codec/%s/raw2real_%d" % (tag, fmt)] = np.stack(outs)
            nbytes = oracle.FORMAT_BYTES[fmt]
            isfloat = fmt >= 8
            full = 1.0 if isfloat else float(2 ** (8 * nbytes - 1))
            y = (rng.uniform(-1.2, 1.2, 2 * L) * full).astype(dt)
            if not isfloat:
                top = full - 128 if (nbytes == 4 and rs == 4) else full - 1
                y[:12] = np.array([0.0, -0.0, 0.49, 0.5, -0.49, -0.5, -0.51, 3.0, -3.0, top, -full, -full + 0.4], dtype=dt)
            G["codec/%s/real_in_%d" % (tag, fmt)] = y
            ov = oracle.Overflow()
            ov.max = 1.0 if isfloat else full - 1
            out = np.zeros(L * nbytes, dtype=np.uint8)
            cv.cbuf2raw(y, out, fmt, 0, 1, False, 0, ov)
            G["codec/%s/real2raw_%d" % (tag, fmt)] = out
            G["codec/%s/real2raw_overflow_%d" % (tag, fmt)] = np.array(ov.as_tuple(), dtype=np.float64)

        # dither: table, map, three dithered blocks on channel 1
        tab = cv.dither_table()
        G["dither/%s/table_head" % tag] = tab[:4096].copy()
        G["dither/%s/table_size_crc" % tag] = np.array([len(tab), zlib.crc32(tab.tobytes())], dtype=np.int64)
        G["dither/%s/map" % tag] = cv.dither_map()
        ov = oracle.Overflow()
        ov.max = 32767.0
        yin, yout, ptrs = [], [], []
        for blk in range(3):
            y = (rng.uniform(-1.05, 1.05, 2 * L) * 32768).astype(dt)
            out = np.zeros(L * 2, dtype=np.uint8)
            cv.cbuf2raw(y, out, oracle.S16_LE, 0, 1, True, 1, ov)
            yin.append(y); yout.append(out.copy()); ptrs.append(cv.dither_ptr(1))
        G["dither/%s/real_in" % tag] = np.stack(yin)
        G["dither/%s/s16_out" % tag] = np.stack(yout)
        G["dither/%s/ptrs" % tag] = np.array(ptrs)
        G["dither/%s/overflow" % tag] = np.array(ov.as_tuple(), dtype=np.float64)

        # engine: brutefir::run, float I/O and dithered S16 output
        P, C = 3, 2
        fmt = oracle.FLOAT_LE if rs == 4 else oracle.FLOAT64_LE
        hh = [(rng.standard_normal(L * P - 3) * np.exp(-np.arange(L * P - 3) / 40.0)) for _ in range(C)]
        G["engine/%s/h" % tag] = np.stack(hh)
        xin = rng.uniform(-1, 1, (8 * L, C))
        G["engine/%s/x" % tag] = xin
        for name, out_fmt, dith in (("float", fmt, False), ("s16", oracle.S16_LE, False), ("s16_dither", oracle.S16_LE, True)):
            e = oracle.Engine(L, P, rs, C, fmt, out_fmt, 2000, dith, kind="ref")
            assert e.set_coeff(hh, P, 0.9) == 0
            outs = []
            for b in range(8):
                raw = np.ascontiguousarray(xin[b * L:(b + 1) * L].astype(dt)).view(np.uint8).ravel()
                rc, out = e.run(raw)
                assert rc == 0
                outs.append(out.copy())
            G["engine/%s/out_%s" % (tag, name)] = np.stack(outs)
            G["engine/%s/overflow_%s" % (tag, name)] = np.array([e.overflow(c).as_tuple() for c in range(C)], dtype=np.float64)
            if dith:
                G["engine/%s/dither_ptrs" % tag] = np.array([e.dither_ptr(c) for c in range(C)])

        # equalizer (next row N1): 31 ISO bands, seeded gains, taps = 64 * 16 = 1024
        bands = np.array([20, 25, 31.5, 40, 50, 63, 80, 100, 125, 160, 200, 250, 315, 400, 500, 630, 800, 1000, 1250,
                          1600, 2000, 2500, 3150, 4000, 5000, 6300, 8000, 10000, 12500, 16000, 20000], dtype=np.float64)
        mag = np.random.default_rng(7).integers(-120, 121, 31) / 10.0
        phase = np.zeros(31)
        G["eq/bands"], G["eq/mag_db"], G["eq/phase"] = bands, mag, phase
        G["eq/%s/render" % tag] = oracle.equalizer_render(64, 16, rs, 48000, bands, mag, phase, kind="ref")

    G["meta/fft_provider"] = np.frombuffer(oracle.lib("ref").fft_provider(), dtype=np.uint8)
    np.savez_compressed(OUT, **G)
    print("wrote %s: %d arrays, %d bytes" % (OUT, len(G), os.path.getsize(OUT)))


if __name__ == "__main__":
    main()
