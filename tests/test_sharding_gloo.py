"""CPU tests of the multi-GPU host logic (foo-dsp-bfir_b200/sharding.py) with world_size 2 and the
`gloo` backend: stream sharding (no collective) and partition sharding (one sum all-reduce of the
partial spectra per block). The per-rank compute is a stand-in built from the oracle's convolver
entry points, so the test exercises exactly the sharding / reduce / sequencing code the GPU ranks run."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, load_package, decay_filter, white_noise, rel_rms


def test_shard_arithmetic():
    import importlib
    sh = importlib.import_module("foo-dsp-bfir_b200.sharding")
    for n, w in [(4096, 8), (10, 4), (3, 8), (512, 1), (0, 2)]:
        spans = [sh.stream_shard(n, w, r) for r in range(w)]
        assert sum(c for _, c in spans) == n
        pos = 0
        for first, count in spans:
            assert first == pos and count in (n // w, n // w + 1)
            pos += count
    assert sh.partition_shard(512, 8, 3) == (192, 64)      # cfg4: P 512 over 8 GPUs
    with pytest.raises(ValueError):
        sh.stream_shard(4, 2, 2)


class OracleShardEngine:
    """run_partial / run_finish of ONE partition shard, restated with the oracle's entry points in the
    order of brutefir::run (brutefir.cpp:252-334). CPU tensors stand in for device buffers."""

    def __init__(self, oracle, torch, L, P, C, coeffs, begin, count):
        self.o, self.torch, self.L, self.N, self.P, self.C = oracle, torch, L, 2 * L, P, C
        self.begin, self.count = begin, count
        self.cv = oracle.Convolver(L, 8, "port")
        self.H = [self.cv.preprocess_coeff(coeffs[c], P) for c in range(C)]
        self.fdl = np.zeros((C, P, self.N))
        self.prev = np.zeros((C, L))
        self.acc = torch.zeros(C * self.N, dtype=torch.float64)
        self.t = 0

    def run_partial_device(self, d_in):
        x = d_in.numpy().reshape(self.L, self.C)
        acc = self.acc.numpy().reshape(self.C, self.N)
        seen = min(self.t + 1, self.P)
        for c in range(self.C):
            tbuf = np.concatenate([self.prev[c], x[:, c]])
            self.prev[c] = x[:, c]
            self.fdl[c, self.t % self.P] = self.cv.mixnscale([self.cv.time2freq(tbuf)], [1.0], 1)
            a = np.zeros(self.N)
            first = True
            for i in range(self.begin, min(self.begin + self.count, seen)):
                slot = (self.t - i) % self.P
                if first:
                    a = self.cv.convolve(self.fdl[c, slot].copy(), self.H[c][i].copy())
                    first = False
                else:
                    self.cv.convolve_add(self.fdl[c, slot].copy(), self.H[c][i].copy(), a)
            acc[c] = a

    def run_finish_device(self, d_out):
        acc = self.acc.numpy().reshape(self.C, self.N)
        y = d_out.numpy().reshape(self.L, self.C)
        for c in range(self.C):
            y[:, c] = self.cv.freq2time(self.cv.mixnscale([acc[c].copy()], [1.0], 3))[: self.L]
        self.t += 1

    def sync(self):
        return 0


def _worker(rank, world, port, L, P, C, nblocks, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import importlib
    import torch
    import torch.distributed as dist
    import oracle
    sh = importlib.import_module("foo-dsp-bfir_b200.sharding")
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    coeffs = [decay_filter(c, L * P - 5) for c in range(C)]
    begin, count = sh.partition_shard(P, world, rank)
    eng = OracleShardEngine(oracle, torch, L, P, C, coeffs, begin, count)
    drv = sh.PartitionShardedEngine(eng, eng.acc)
    x = white_noise(7, nblocks * L, C)
    outs = []
    for b in range(nblocks):
        d_in = torch.from_numpy(np.ascontiguousarray(x[b * L:(b + 1) * L]).ravel())
        d_out = torch.zeros(L * C, dtype=torch.float64)
        drv.run_device(d_in, d_out)
        outs.append(d_out.numpy().reshape(L, C).copy())
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), np.concatenate(outs))
    # stream sharding: each rank owns whole streams, results must not depend on the world size
    first, cnt = sh.stream_shard(5, world, rank)
    np.save(os.path.join(out_dir, "streams%d.npy" % rank), np.arange(first, first + cnt))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_partition_sharding_world2_gloo(tmp_path, oracle):
    import torch.multiprocessing as mp
    L, P, C, nblocks, world = 64, 5, 2, 9, 2
    mp.spawn(_worker, args=(world, _free_port(), L, P, C, nblocks, str(tmp_path)), nprocs=world, join=True)
    y0 = np.load(tmp_path / "rank0.npy")
    y1 = np.load(tmp_path / "rank1.npy")
    assert np.array_equal(y0, y1)                      # every rank holds the full reduced result
    ref = oracle.Engine(L, P, 8, C, oracle.FLOAT64_LE, oracle.FLOAT64_LE, 44100, False, kind="port")
    assert ref.set_coeff([decay_filter(c, L * P - 5) for c in range(C)], P) == 0
    x = white_noise(7, nblocks * L, C)
    ys = []
    for b in range(nblocks):
        rc, out = ref.run(np.ascontiguousarray(x[b * L:(b + 1) * L]).view(np.uint8).ravel())
        ys.append(out.view(np.float64).reshape(L, C))
    assert rel_rms(y0, np.concatenate(ys)) < 1e-13     # shard sum order differs from the serial loop
    s0, s1 = np.load(tmp_path / "streams0.npy"), np.load(tmp_path / "streams1.npy")
    assert list(np.concatenate([s0, s1])) == [0, 1, 2, 3, 4]
