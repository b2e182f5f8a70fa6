"""Fused partition-shard reduce (peer stores instead of an NCCL reduce) emulated on ONE GPU: two engines
of one process play ranks 0 and 1 of a world of 2 on one stream and are connected with
bfir_peer_set_ptr, so the producing kernels store their partial spectra into the other engine's receive
buffer exactly as they would over NVLink. (Kernels that wait on each other are never launched; the
cross-rank barrier of the real multi-process run is stream order here.)"""
import numpy as np
import pytest

from conftest import white_noise, decay_filter, rel_rms

pytestmark = pytest.mark.gpu


def build(pkg, L, P, rs, C, fmt, xb, h, gains, **kw):
    e = pkg.Brutefir(L, P, rs, C, fmt, fmt, 48000, False, n_groups=1, xbar_inputs=xb[0], xbar_outputs=xb[1], **kw)
    assert e.set_coeff(h, P) == 0
    if xb[0]:
        e.set_crossbar(*gains)
    return e


@pytest.mark.parametrize("rs,L,P,C,xb", [(4, 512, 6, 4, (0, 0)), (8, 256, 5, 5, (0, 0)), (4, 1024, 4, 4, (3, 5)), (4, 32768, 4, 8, (8, 8))])
def test_two_emulated_ranks_match_unsharded(pkg, rs, L, P, C, xb):
    import torch
    fmt = pkg.FLOAT_LE if rs == 4 else pkg.FLOAT64_LE
    dt, tdt = (np.float32, torch.float32) if rs == 4 else (np.float64, torch.float64)
    n_in = xb[0] or C
    n_out = xb[1] or C
    rng = np.random.default_rng(C)
    gains = (rng.standard_normal((C, n_in)) / np.sqrt(n_in), rng.standard_normal((n_out, C)) / np.sqrt(C))
    h = [decay_filter(c, L * P) for c in range(C)]
    full = build(pkg, L, P, rs, C, fmt, xb, h, gains)
    world = 2
    ranks = []
    for r in range(world):
        begin = r * (P // world)
        count = P // world if r < world - 1 else P - begin
        e = build(pkg, L, P, rs, C, fmt, xb, h, gains, part_begin=begin, part_count=count)
        e.peer_setup(r, world)
        ranks.append(e)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    for e in ranks + [full]:
        e.set_stream(stream.cuda_stream)
    for r, e in enumerate(ranks):
        for q, other in enumerate(ranks):
            if q != r:
                e.peer_set_ptr(q, other.peer_recv_ptr())
    own = [e.peer_own_channels() for e in ranks]
    assert own[0][0] == 0 and own[1][0] == own[0][1] and own[0][1] + own[1][1] == n_out
    nb = 2 * P + 1
    x = white_noise(1, nb * L, n_in).astype(dt)
    d_ref = torch.empty(L * n_out, dtype=tdt, device="cuda")
    d_own = [torch.empty(L * c, dtype=tdt, device="cuda") for _, c in own]
    for b in range(nb):
        d_in = torch.from_numpy(np.ascontiguousarray(x[b * L:(b + 1) * L]).ravel()).cuda()
        full.run_device(d_in, d_ref)
        for e in ranks:
            e.run_partial_device(d_in)          # pushes into the peers' receive buffers
        for e, o in zip(ranks, d_own):
            e.run_finish_device(o)              # (same stream: every push is complete here)
        for e in ranks + [full]:
            assert e.sync() == 0
        ref = d_ref.cpu().numpy().reshape(L, n_out)
        got = np.concatenate([o.cpu().numpy().reshape(L, c) for o, (_, c) in zip(d_own, own)], axis=1)
        assert rel_rms(got, ref) < (2e-6 if rs == 4 else 1e-13), b
    assert [e.blockcounter() for e in ranks] == [nb, nb]


@pytest.mark.parametrize("rs,L,P,C,xb", [(4, 512, 6, 4, (0, 0)), (8, 256, 4, 5, (0, 0)), (4, 1024, 8, 4, (3, 5)), (4, 2048, 16, 8, (8, 8))])
def test_two_emulated_ranks_four_blocks_per_call(pkg, rs, L, P, C, xb):
    """bfir_run_partial_quad_device / bfir_run_finish_quad_device: four blocks per call on partition shards (one
    four-block partition sum over the rank's partitions, partial results pushed into phases 2..9 of the owners'
    receive buffers, arrival flags instead of a collective), after a one-block prefill, mixed with one-block calls:
    same output as the unsharded engine. One-block calls before the delay line is full are refused with NOT_READY."""
    import torch
    fmt = pkg.FLOAT_LE if rs == 4 else pkg.FLOAT64_LE
    dt, tdt = (np.float32, torch.float32) if rs == 4 else (np.float64, torch.float64)
    n_in = xb[0] or C
    n_out = xb[1] or C
    rng = np.random.default_rng(C + 1)
    gains = (rng.standard_normal((C, n_in)) / np.sqrt(n_in), rng.standard_normal((n_out, C)) / np.sqrt(C))
    h = [decay_filter(c, L * P) for c in range(C)]
    full = build(pkg, L, P, rs, C, fmt, xb, h, gains)
    world = 2
    ranks = []
    for r in range(world):
        begin = r * (P // world)
        count = P // world if r < world - 1 else P - begin
        e = build(pkg, L, P, rs, C, fmt, xb, h, gains, part_begin=begin, part_count=count)
        e.peer_setup(r, world)
        ranks.append(e)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    for e in ranks + [full]:
        e.set_stream(stream.cuda_stream)
    for r, e in enumerate(ranks):
        for q, other in enumerate(ranks):
            if q != r:
                e.peer_set_ptr(q, other.peer_recv_ptr())
    own = [e.peer_own_channels() for e in ranks]
    nb = P + 4 * 5 + 2
    x = white_noise(2, nb * L, n_in).astype(dt)
    d_in = [torch.from_numpy(np.ascontiguousarray(x[b * L:(b + 1) * L]).ravel()).cuda() for b in range(nb)]
    d_ref = [torch.empty(L * n_out, dtype=tdt, device="cuda") for _ in range(nb)]
    d_own = [[torch.empty(L * c, dtype=tdt, device="cuda") for _ in range(nb)] for _, c in own]
    torch.cuda.synchronize()
    for b in range(nb):
        full.run_device(d_in[b], d_ref[b])
    with pytest.raises(pkg.BfirError):
        ranks[0].run_partial_quad_device(d_in[0:4])           # delay line not filled yet

    def single(b):
        for e in ranks:
            e.run_partial_device(d_in[b])
        for e, o in zip(ranks, d_own):
            e.run_finish_device(o[b])

    b = 0
    while b < P:
        single(b)
        b += 1
    for call in range(5):
        if call == 3:                                         # a one-block call between the four-block ones
            single(b)
            b += 1
        for e in ranks:
            e.run_partial_quad_device(d_in[b:b + 4])
        for e, o in zip(ranks, d_own):
            e.run_finish_quad_device(o[b:b + 4])
        b += 4
    while b < nb:
        single(b)
        b += 1
    for e in ranks + [full]:
        assert e.sync() == 0
    for b in range(nb):
        ref = d_ref[b].cpu().numpy().reshape(L, n_out)
        got = np.concatenate([o[b].cpu().numpy().reshape(L, c) for o, (_, c) in zip(d_own, own)], axis=1)
        assert rel_rms(got, ref) < (2e-6 if rs == 4 else 1e-13), b
    assert [e.blockcounter() for e in ranks] == [nb, nb]


@pytest.mark.parametrize("rs,L,P,C,xb,shard_inputs", [(4, 512, 6, 4, (0, 0), 0), (8, 256, 4, 5, (0, 0), 0), (4, 1024, 8, 4, (3, 5), 0),
                                                      (4, 2048, 16, 8, (8, 8), 0), (4, 1024, 8, 4, (3, 5), 1), (4, 2048, 16, 8, (8, 8), 1),
                                                      (8, 512, 5, 6, (4, 4), 1), (4, 16384, 4, 12, (0, 0), 0), (8, 8192, 3, 12, (6, 6), 0), (8, 8192, 3, 12, (6, 6), 1)])
def test_two_emulated_ranks_four_blocks_staged(pkg, rs, L, P, C, xb, shard_inputs, monkeypatch):
    """bfir_run_shard_quad_staged: the four-block shard call through the stage pipeline (forward transforms, partition
    sum + pushes, arrival wait + output stage of neighbouring calls on three streams per rank; receive-buffer phases
    handed over by the arrival flags alone). Seven staged calls with a one-block call and a split (partial / finish)
    four-block call in between: same output as the unsharded engine for every block. shard_inputs = 1: with a crossbar
    every rank transforms only its own inputs and stores the spectra into its peers' input regions (second flag array;
    BFIR_SHARD_INPUTS=1, off by default because it measured slower), including the hand-back of the previous-block rows
    to the one-block calls in between."""
    import torch
    monkeypatch.setenv("BFIR_SHARD_INPUTS", str(shard_inputs))
    fmt = pkg.FLOAT_LE if rs == 4 else pkg.FLOAT64_LE
    dt, tdt = (np.float32, torch.float32) if rs == 4 else (np.float64, torch.float64)
    n_in = xb[0] or C
    n_out = xb[1] or C
    rng = np.random.default_rng(C + 1)
    gains = (rng.standard_normal((C, n_in)) / np.sqrt(n_in), rng.standard_normal((n_out, C)) / np.sqrt(C))
    h = [decay_filter(c, L * P) for c in range(C)]
    full = build(pkg, L, P, rs, C, fmt, xb, h, gains)
    world = 2
    ranks = []
    for r in range(world):
        begin = r * (P // world)
        count = P // world if r < world - 1 else P - begin
        e = build(pkg, L, P, rs, C, fmt, xb, h, gains, part_begin=begin, part_count=count)
        e.peer_setup(r, world)
        ranks.append(e)
    # every emulated rank keeps its OWN streams, as separate processes would: with the sharded input stage a rank's
    # forward stream waits for its peers' input flags, which must not sit behind this rank's work on a shared stream
    for r, e in enumerate(ranks):
        for q, other in enumerate(ranks):
            if q != r:
                e.peer_set_ptr(q, other.peer_recv_ptr())
    own = [e.peer_own_channels() for e in ranks]
    nb = P + 4 * 8 + 1 + 8 * 3
    x = white_noise(3, nb * L, n_in).astype(dt)
    d_in = [torch.from_numpy(np.ascontiguousarray(x[b * L:(b + 1) * L]).ravel()).cuda() for b in range(nb)]
    d_ref = [torch.empty(L * n_out, dtype=tdt, device="cuda") for _ in range(nb)]
    d_own = [[torch.empty(L * c, dtype=tdt, device="cuda") for _ in range(nb)] for _, c in own]
    torch.cuda.synchronize()
    for b in range(nb):
        full.run_device(d_in[b], d_ref[b])

    def single(b):
        for e in ranks:
            e.run_partial_device(d_in[b])
        torch.cuda.synchronize()                              # the cross-rank barrier of the one-block calls
        for e, o in zip(ranks, d_own):
            e.run_finish_device(o[b])
        torch.cuda.synchronize()

    b = 0
    while b < P:
        single(b)
        b += 1
    torch.cuda.synchronize()                                  # staged contract: inputs complete (they are), prefill done
    for call in range(8):
        if call == 3:                                         # a one-block call closes the pipeline, the next call reopens it
            single(b)
            b += 1
        if call == 5:                                         # the split four-block call between staged ones
            for e in ranks:
                e.run_partial_quad_device(d_in[b:b + 4])
            for e, o in zip(ranks, d_own):
                e.run_finish_quad_device(o[b:b + 4])
        else:
            for e, o in zip(ranks, d_own):
                e.run_shard_quad_staged(d_in[b:b + 4], o[b:b + 4])
        b += 4
    for call in range(3):                                     # eight blocks per call (two four-block calls where the shard is too small)
        for e, o in zip(ranks, d_own):
            e.run_shard_oct_staged(d_in[b:b + 8], o[b:b + 8])
        b += 8
    assert b == nb
    for e in ranks + [full]:
        assert e.sync() == 0
    for b in range(nb):
        ref = d_ref[b].cpu().numpy().reshape(L, n_out)
        got = np.concatenate([o[b].cpu().numpy().reshape(L, c) for o, (_, c) in zip(d_own, own)], axis=1)
        assert rel_rms(got, ref) < (2e-6 if rs == 4 else 1e-13), b
    assert [e.blockcounter() for e in ranks] == [nb, nb]


def test_unconnected_peer_is_refused(pkg):
    e = pkg.Brutefir(256, 4, 4, 4, pkg.FLOAT_LE, pkg.FLOAT_LE, 48000, False, part_begin=0, part_count=2)
    e.set_coeff([decay_filter(c, 1024) for c in range(4)], 4)
    e.peer_setup(0, 2)
    import torch
    d = torch.zeros(256 * 4, dtype=torch.float32, device="cuda")
    with pytest.raises(pkg.BfirError):
        e.run_partial_device(d)
    with pytest.raises(pkg.BfirError):
        pkg.Brutefir(256, 4, 4, 2, pkg.FLOAT_LE, pkg.FLOAT_LE, 48000, False).peer_setup(0, 3)   # world > channels
