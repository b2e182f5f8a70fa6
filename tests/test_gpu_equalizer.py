"""GPU parity of the device equalizer (next row N1: equalizer::generate + render_f/d, reference
brutefir/equalizer.cpp:87-140, 212-394) against the oracle (the unmodified reference render where
oracle/_ref exists, else the port) and the golden vectors, plus the BASELINE configs[2] flow:
equalizer-generated coefficients swapped in with a crossfade on every block."""
import os

import numpy as np
import pytest

from conftest import rel_rms, white_noise, parity, SwapChain

pytestmark = pytest.mark.gpu
BANDS = [20, 25, 31.5, 40, 50, 63, 80, 100, 125, 160, 200, 250, 315, 400, 500, 630, 800, 1000, 1250, 1600, 2000, 2500,
         3150, 4000, 5000, 6300, 8000, 10000, 12500, 16000, 20000]
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz")


def gains(seed):
    return np.random.default_rng(seed).integers(-120, 121, 31) / 10.0


@pytest.mark.parametrize("block,n_blocks,rs,rate", [(64, 16, 4, 48000), (64, 16, 8, 48000), (1024, 64, 8, 44100),
                                                    (4096, 64, 4, 96000), (4096, 64, 8, 96000), (16, 2, 8, 44100), (1024, 8, 4, 32000)])
def test_render_matches_oracle(pkg, oracle, block, n_blocks, rs, rate):
    """(1024, 64, 8) is the product's equalizer (common.h:17-19); (4096, 64) is configs[2]: 262144 taps ->
    131072 coefficients, a four-step inverse transform"""
    eq = pkg.Equalizer(block, n_blocks, rs, rate)
    assert eq.taps == block * n_blocks
    mag = gains(block + rs)
    phase = np.random.default_rng(3).uniform(-90, 90, 31)
    got = eq.generate(BANDS, mag, phase)
    ref = oracle.equalizer_render(block, n_blocks, rs, rate, BANDS, mag, phase)
    assert got.shape == ref.shape == (block * n_blocks // 2,)
    # float: rad = -pi*n is formed in float32 (equalizer.cpp:251), so cosf sees arguments up to 4e5 and a
    # 1-ulp difference between CUDA's and glibc's cosf/sinf is amplified; 1e-5 relative RMS still holds
    assert rel_rms(got, ref) < (1e-5 if rs == 4 else 1e-12)
    # flat 0 dB, zero phase -> a unit impulse in the middle of the filter (linear-phase centring)
    flat = eq.generate(BANDS, np.zeros(31), np.zeros(31))
    want = np.zeros_like(flat)
    want[0] = 1.0            # upper half of the taps-point frame: centre tap taps/2 is its first sample
    # (float32: the reference forms rad = -pi*n in float, good to ~0.03 rad at n = 1e5, so its own float
    #  filters are only accurate to ~1e-2 at this length; the double build is exact)
    assert np.max(np.abs(flat - want)) < (2e-2 if rs == 4 else 1e-9)


def test_render_matches_golden(pkg):
    G = np.load(GOLDEN)
    for rs, tag in ((4, "f32"), (8, "f64")):
        eq = pkg.Equalizer(64, 16, rs, 48000)
        got = eq.generate(G["eq/bands"], G["eq/mag_db"], G["eq/phase"])
        assert rel_rms(got, G["eq/%s/render" % tag]) < (1e-5 if rs == 4 else 1e-12)


def test_fewer_bands_and_invalid_sizes(pkg, oracle):
    eq = pkg.Equalizer(256, 4, 8, 44100)
    sel = [3, 10, 17, 24]
    f = [BANDS[i] for i in sel]
    m = [6.0, -3.0, 4.5, -9.0]
    p = [0.0, 10.0, -20.0, 0.0]
    assert rel_rms(eq.generate(f, m, p), oracle.equalizer_render(256, 4, 8, 44100, f, m, p)) < 1e-12
    with pytest.raises(pkg.BfirError):
        pkg.Equalizer(100, 3, 8, 44100)          # "Equalizer length is not a power of two" (equalizer.cpp:38-42)
    with pytest.raises(pkg.BfirError):
        eq.generate(list(range(1, 33)), [0.0] * 32, [0.0] * 32)   # more than BAND_COUNT bands (:96-100)


def _cfg2_flow(pkg, oracle, L, EQB, C, nb, tag):
    """configs[2]: a fresh equalizer curve every block, rendered on the device, handed to the engine without leaving
    the GPU (bfir_set_coeff_device, stride 0 = same filter for every channel) and cross-faded in on EVERY block; the
    oracle renders the same curves on the CPU (the reference's render_f) and composes run() + crossfade_inplace
    (conftest.SwapChain); the same chain in double, fed the reference's float coefficients, is the float64 truth."""
    rs = 4
    taps = L * EQB
    P = (taps // 2) // L
    eq = pkg.Equalizer(L, EQB, rs, 96000)
    g = pkg.Brutefir(L, P, rs, C, pkg.FLOAT_LE, pkg.FLOAT_LE, 96000, False)
    ref = SwapChain(oracle.Convolver(L, rs), L, P, C, np.float32)
    tru = SwapChain(oracle.Convolver(L, 8, kind="port"), L, P, C, np.float64)
    mags = [gains(7 + b) for b in range(nb)]
    zero = np.zeros(31)
    assert g.set_coeff_device(eq.generate_device(BANDS, mags[0], zero), 0, C, taps // 2, P) == 0
    x = white_noise(5, nb * L, C).astype(np.float32)
    got, want, truth = [], [], []
    H_prev = H64_prev = None
    for t in range(nb):
        h_ref = oracle.equalizer_render(L, EQB, rs, 96000, BANDS, mags[t], zero)
        H, H64 = ref.spectra([h_ref], same_for_all=True), tru.spectra([h_ref], same_for_all=True)
        if t >= 1:
            assert g.set_coeff_device(eq.generate_device(BANDS, mags[t], zero), 0, C, taps // 2, P, crossfade=True) == 0
        blk = np.ascontiguousarray(x[t * L:(t + 1) * L])
        rc, out = g.run(blk.view(np.uint8).ravel())
        assert rc == 0
        got.append(out.view(np.float32).reshape(L, C).copy())
        want.append(ref.block(blk, H, H_prev))
        truth.append(tru.block(blk, H64, H64_prev))
        H_prev, H64_prev = H, H64
    got, want, truth = np.concatenate(got), np.concatenate(want), np.concatenate(truth)
    for c in range(C):
        parity("%s/ch%d" % (tag, c), got[:, c], want[:, c], 1e-5, truth[:, c])
        worst = max(rel_rms(got[b * L:(b + 1) * L, c], want[b * L:(b + 1) * L, c]) for b in range(nb))
        print("%s ch%d worst single block vs oracle: %.3e" % (tag, c, worst))


def test_cfg2_equalizer_crossfade_every_block(pkg, oracle):
    """configs[2] at reduced size: 4096-tap equalizer -> 2048 coefficients -> P 8, L 256"""
    _cfg2_flow(pkg, oracle, 256, 16, 2, 18, "cfg2_small")


def test_cfg2_baseline_size_equalizer_crossfade_every_block(pkg, oracle):
    """configs[2] at its BASELINE size: stereo 96 kHz float, L 4096, 262144-point equalizer render (four-step inverse
    transform on the device) -> 131072 coefficients -> P 32, a crossfade swap on every one of 44 blocks (the last 12
    with all 32 partitions live)"""
    _cfg2_flow(pkg, oracle, 4096, 64, 2, 44, "cfg2_baseline_size")
