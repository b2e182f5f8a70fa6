"""Multi-GPU host logic: one process per GPU, `torch.distributed` for the plumbing.

Two ways the path shards, and nothing else:

* **streams / channels** -- channels never interact in ``brutefir::run`` (filter n <-> channel n only,
  reference brutefir/brutefir.cpp:213-216, 252-334), so whole streams are dealt to ranks with
  :func:`stream_shard` and there is NO data-path collective.
* **partitions of one long filter** -- the partition sum ``sum_i X[t-i] * H[i]`` (brutefir.cpp:288-299)
  is associative across ``i``: rank g convolves partitions :func:`partition_shard` of every channel,
  every rank transforms the (small) input block redundantly, and ONE sum-reduction of the partial
  output spectra (``channels * 2L`` reals) per block joins them (:class:`PartitionShardedEngine`).

The engine object only has to provide ``run_partial_device``, ``run_finish_device``, ``sync`` and a
tensor view of its partial spectra, so the same driver runs on CPU tensors with the ``gloo`` backend in
the tests (tests/test_sharding_gloo.py) and on the CUDA engine with ``nccl`` on the GPUs.
"""


def stream_shard(n_streams, world_size, rank):
    """Contiguous split of `n_streams` independent streams: returns (first_stream, count) of `rank`.
    The first ``n_streams % world_size`` ranks get one stream more."""
    if world_size < 1 or not 0 <= rank < world_size or n_streams < 0:
        raise ValueError("invalid shard request")
    base, extra = divmod(n_streams, world_size)
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def partition_shard(n_partitions, world_size, rank):
    """Contiguous split of the filter partitions [0, P): returns (part_begin, part_count) of `rank`.
    Ranks beyond the partition count get an empty shard (count 0) and contribute zeros to the reduce."""
    return stream_shard(n_partitions, world_size, rank)


class PartitionShardedEngine:
    """Drives one partition shard per rank and joins the shards with a sum all-reduce.

    engine      object with run_partial_device(d_in), run_finish_device(d_out), sync()
    acc         tensor view (torch) of the engine's partial accumulated spectra, reduced in place
    group       torch.distributed process group (None = default)
    """

    def __init__(self, engine, acc, group=None):
        import torch.distributed as dist
        self.dist, self.engine, self.acc, self.group = dist, engine, acc, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.blocks = 0

    def run_device(self, d_in, d_out, pre_reduce=None):
        """One block: partial partition sum on every rank -> all-reduce(sum) -> output stage.
        `pre_reduce` is called between the partial step and the collective (the CUDA engine makes the
        communication stream wait for the compute stream there)."""
        self.engine.run_partial_device(d_in)
        if pre_reduce is not None:
            pre_reduce()
        if self.world > 1:
            self.dist.all_reduce(self.acc, op=self.dist.ReduceOp.SUM, group=self.group)
        self.engine.run_finish_device(d_out)
        self.blocks += 1

    def sync(self):
        return self.engine.sync()


def make_partition_sharded(pkg, filter_length, filter_blocks, realsize, channels, in_format, out_format,
                           sampling_rate, apply_dither, coeffs, coeff_blocks=None, scale=1.0, n_streams=1,
                           device=-1, group=None, xbar_inputs=0, xbar_outputs=0, in_gains=None, out_gains=None):
    """Build this rank's shard of a partition-sharded CUDA engine and its driver (NCCL).

    Every rank passes the FULL coefficient arrays; the shard convolves only its partitions. The engine
    runs on torch's current CUDA stream so the NCCL all-reduce is ordered with the kernels."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    begin, count = partition_shard(filter_blocks, world, rank)
    if count == 0:
        raise ValueError("more ranks than partitions: %d > %d" % (world, filter_blocks))
    eng = pkg.Brutefir(filter_length, filter_blocks, realsize, channels, in_format, out_format, sampling_rate,
                       apply_dither, n_streams=n_streams, device=device, part_begin=begin, part_count=count,
                       n_groups=1, xbar_inputs=xbar_inputs, xbar_outputs=xbar_outputs)
    if xbar_inputs or xbar_outputs:
        eng.set_crossbar(in_gains, out_gains)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    rc = eng.set_coeff(coeffs, filter_blocks if coeff_blocks is None else coeff_blocks, scale)
    if rc != 0:
        raise RuntimeError("set_coeff failed: %d" % rc)
    ptr, nbytes = eng.acc_device_ptr()
    acc = pkg.as_torch(ptr, nbytes // realsize, "<f4" if realsize == 4 else "<f8")
    return PartitionShardedEngine(eng, acc, group), eng


class FusedPartitionShardedEngine:
    """Partition sharding with the reduce fused into the producing kernel: every rank's partition-sum
    (or, with a crossbar, output-mix) kernel stores its partial spectra straight into the receive buffer
    of the rank that owns the channel (peer-mapped memory, NVLink stores), a one-element all-reduce on
    the same stream is the only collective (it orders "all peers have pushed" before "owners sum"), and
    every rank then emits its own channels as a compact interleaved block [L][own_count]."""

    def __init__(self, engine, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.engine, self.group = torch, dist, engine, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        engine.peer_setup(self.rank, self.world)
        handles = [None] * self.world
        dist.all_gather_object(handles, engine.peer_export(), group=group)
        for q, h in enumerate(handles):
            engine.peer_import(q, h)
        self.own_first, self.own_count = engine.peer_own_channels()
        self.flag = torch.zeros(1, dtype=torch.float32, device="cuda")
        dist.barrier(group=group)

    def run_device(self, d_in, d_out_own):
        self.engine.run_partial_device(d_in)                                  # pushes partials to the owners
        self.dist.all_reduce(self.flag, op=self.dist.ReduceOp.SUM, group=self.group)   # cross-rank barrier on the stream
        self.engine.run_finish_device(d_out_own)                              # sum own slots + output stage

    def run_device_quad(self, d_ins, d_outs_own):
        """Four blocks per call (steady state, i.e. after filter_blocks one-block calls): one four-block partition
        sum per rank, pushes, and a device-side arrival flag instead of the collective -- no NCCL call at all."""
        self.engine.run_partial_quad_device(d_ins)
        self.engine.run_finish_quad_device(d_outs_own)

    def run_device_quad_staged(self, d_ins, d_outs_own):
        """The same through the engine's stage pipeline (bfir_run_shard_quad_staged): forward transforms of the next
        call and the wait + sum + output stage of the previous one run beside the partition sum on side streams.
        The inputs must be complete when the call is made; call join() (stream-ordered) or sync() before reading
        the outputs."""
        self.engine.run_shard_quad_staged(d_ins, d_outs_own)

    def run_device_oct_staged(self, d_ins, d_outs_own):
        """Eight blocks per call through the stage pipeline (bfir_run_shard_oct_staged)."""
        self.engine.run_shard_oct_staged(d_ins, d_outs_own)

    def join(self):
        return self.engine.join()

    def gather(self, d_out_own, d_all):
        """d_all [world][L * cpr] <- every rank's own block (only needed when one rank wants all channels)"""
        self.dist.all_gather_into_tensor(d_all, d_out_own, group=self.group)

    def sync(self):
        return self.engine.sync()
