#!/usr/bin/env python
"""bench.py -- headline benchmark of the partitioned-convolution hot path (BASELINE.json metric:
"Msamples/s (all channels) at 262144 taps", configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--streams S] [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (one "step" = one block of the block loop, brutefir::run, for every stream of the batch):
  cfg1: 7.1 (8 channels) @ 48 kHz, double precision (realsize 8), FLOAT64_LE in/out, 262144-tap FIR per
  channel = P 32 partitions of L 8192 samples (FFT 16384 points), distinct filter per channel.
  S independent 7.1 streams are batched per GPU so that the streamed set (coefficient spectra + delay
  line = S x 64 MiB) exceeds the 126 MB L2 -- no L2 flush is needed between steps. N GPUs: every rank
  runs its own S streams (channel/stream sharding, no collective): weak scaling.

Printed JSON (one line, rank 0): value = whole-job Msamples/s with inputs resident in HBM, eight blocks per call
(bfir_run_device_oct, staged: ONE partition-sum launch for the eight, every coefficient spectrum read once) through the
engine's stage pipeline (transforms of the neighbouring calls on side streams beside the partition sum); e2e = the
same metric on pinned HOST buffers, the H2D of every input block and the D2H of every output block inside
the timed region: e2e.value through bfir_run_async_pair/bfir_wait (three calls in flight), e2e.one_block_per_call
through bfir_run_async, e2e.sync_run through the reference's synchronous run() = bfir_run (H2D + kernels + D2H +
sync per call); roofline = the eight-block partition-sum kernel, bytes that must move ((2P+15) N realsize per channel
and launch) over its CUDA-event time, with the four-, two- and one-block kernels beside it; cpu_baseline = the
reference's own sources (oracle/_ref, FFT provider named) on the host cores; latency = host-visible bfir_run latency
of ONE 7.1 stream (p50/p99).
`--impl reference` times only the CPU reference (rank 0), same metric/config, --steps / --warmup honoured (each step a
bounded sample: one block of one stream per host thread).
`configs` (unless --no-configs): the other BASELINE.json configurations -- cfg0 / cfg2 block latency and the dither
kernel (rank 0, one GPU), cfg3 (4096 stereo streams, stream-sharded over the ranks, quad kernel + its roofline),
`partition_sharded`: cfg4 (16 Mi taps, 32x32 crossbar) with its partitions sharded over the ranks and the spectrum
reduce (NCCL and fused), against the unsharded engine on rank 0 (tools/bench_configs.py).
"""
import argparse
import importlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CFG = dict(name="cfg1", channels=8, realsize=8, L=8192, P=32, rate=48000, fmt=10)  # FLOAT64_LE
METRIC = "Msamples/s (all channels) at 262144 taps"


def workload_config(streams, n_gpus):
    return {
        "workload": "cfg1: 7.1 room correction 48 kHz, 262144-tap double-precision FIR per channel, "
                    "8192-sample partitions (P=32, FFT 16384), FLOAT64_LE in/out, distinct filters",
        "streams_per_gpu": streams, "channels_per_stream": CFG["channels"], "block": CFG["L"], "partitions": CFG["P"],
        "samples_per_step": streams * CFG["channels"] * CFG["L"] * n_gpus,
        "parallelism": "stream-sharded x%d (no collective)" % n_gpus,
        "l2": "streamed set per step (%d MiB coefficient + delay-line spectra per GPU) exceeds the 126 MB L2; no flush"
              % (streams * CFG["channels"] * 2 * CFG["P"] * 2 * CFG["L"] * CFG["realsize"] // (1 << 20)),
        "prefill_blocks": CFG["P"],
        "step": "one block (8192 frames) of every stream; `value` calls the eight-block entry point (bfir_run_device_oct, staged: eight steps per call "
                "with ONE partition-sum launch), e2e the two-block one on pinned host buffers (bfir_run_async_pair); the four-, two- and one-block-per-call "
                "device numbers ride beside them under `roofline`",
    }


def make_filters(n_channels, taps, first=0):
    out = []
    for ch in range(n_channels):
        g = np.random.default_rng(1000 + first + ch).standard_normal(taps)
        h = g * np.exp(-6.9 * np.arange(taps) / taps)
        out.append(h / np.sqrt(np.sum(h * h)))
    return out


def noise_block(seed, streams, L, C):
    return np.random.default_rng(0xB200 + seed).uniform(-1.0, 1.0, size=(streams, L, C))


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self.stop_flag = threading.Event()
        self.busy = threading.Event()

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {}
            for n in dir(pynvml):
                if n.startswith("nvmlClocksEventReason") or n.startswith("nvmlClocksThrottleReason"):
                    v = getattr(pynvml, n)
                    if isinstance(v, int) and v not in (0,):
                        names.setdefault(v, n.replace("nvmlClocksEventReason", "").replace("nvmlClocksThrottleReason", ""))
            while not self.stop_flag.is_set():
                if self.busy.is_set():
                    self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                    try:
                        r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                    except Exception:
                        r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for bit, n in names.items():
                        if r & bit and bit & (bit - 1) == 0:
                            self.reasons.add(n)
                time.sleep(0.02)
        except Exception as e:  # no NVML: report it instead of failing the bench
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)

    def summary(self):
        s = sorted(self.samples)
        reasons = sorted(x for x in self.reasons if x.lower() not in ("gpuidle", "none", "applicationsclockssetting"))
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "samples": len(s), "reasons": reasons}


# ------------------------------------------------------------------------------------------ CPU reference
def cpu_reference(n_threads, blocks, steps, warmup):
    """The reference's own brutefir::run (oracle/_ref = unmodified sources, on the fastest FFT provider this image
    has: MKL DFTI when libtorch_cpu.so loads, else oracle/fft_r2r; else the port) on the host cores: one engine
    instance (= one 7.1 stream, single-threaded like the reference) per thread. One step = `blocks` block(s) of every
    thread's stream; W untimed steps, then exactly `steps` timed ones."""
    import oracle
    kind = oracle.best_timing_kind()
    C, L, P, rs, fmt, rate = CFG["channels"], CFG["L"], CFG["P"], CFG["realsize"], CFG["fmt"], CFG["rate"]
    engines, inputs = [], []
    for t in range(n_threads):
        e = oracle.Engine(L, P, rs, C, fmt, fmt, rate, False, kind=kind)
        assert e.set_coeff(make_filters(C, L * P, first=t * C), P) == 0
        engines.append(e)
        inputs.append([np.ascontiguousarray(noise_block(100 * t + b, 1, L, C)[0]).view(np.uint8).ravel() for b in range(4)])
    outs = [np.zeros(L * C * 8, dtype=np.uint8) for _ in range(n_threads)]

    def work(t, n):
        for b in range(n):
            rc, _ = engines[t].run(inputs[t][b % 4], outs[t])
            assert rc == 0

    def parallel(n):
        th = [threading.Thread(target=work, args=(t, n)) for t in range(n_threads)]
        t0 = time.perf_counter()
        for x in th:
            x.start()
        for x in th:
            x.join()
        return time.perf_counter() - t0

    parallel(P)                      # prefill: all P partitions active (procblocks == P)
    for _ in range(warmup):
        parallel(blocks)
    times = [parallel(blocks) for _ in range(steps)]
    total = sum(times)
    samples = n_threads * C * L * blocks * steps
    return {
        "value": samples / total / 1e6, "unit": "Msamples/s", "cores": n_threads,
        "kind": "port" if kind == "port" else "reference",
        "sample": "%d thread(s) x 1 stream (8 ch) each, %d timed block(s) of 8192 frames per step x %d steps after a %d-block prefill "
                  "and %d warm-up step(s); FFT provider: %s" % (n_threads, blocks, steps, P, warmup, oracle.lib(kind).fft_provider().decode()),
        "ms_per_step": 1e3 * total / steps, "samples_per_step": n_threads * C * L * blocks,
    }


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


# ------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--streams", type=int, default=16, help="7.1 streams batched per GPU")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the cfg0 / cfg2 / cfg3 / cfg4 blocks")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = max(args.gpus, world)
    K, W, S = args.steps, max(args.warmup, 0), args.streams
    C, L, P, rs, fmt, rate = CFG["channels"], CFG["L"], CFG["P"], CFG["realsize"], CFG["fmt"], CFG["rate"]

    if args.impl == "reference":
        if rank != 0:
            return 0
        cores = host_cores()
        # the timed sample per step is ONE block of one 7.1 stream per host thread (the GPU arm's step is one block of
        # 16 streams per GPU); value is samples per second, so the two arms compare directly. K and W are honoured.
        r = cpu_reference(cores, blocks=1, steps=max(1, K), warmup=W)
        line = {
            "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "Msamples/s", "n_gpus": n_gpus,
            "steps": max(1, K), "warmup": W, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(S, n_gpus),
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pkg = importlib.import_module("foo-dsp-bfir_b200")
    pkg.load_library()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local_rank)
    sampler.start()

    # n_groups=1 for the device-resident pass: kernels run back to back on one stream, so the CUDA-event
    # time of each kernel (roofline) is its own; the end-to-end pass below switches group pipelining on
    eng = pkg.Brutefir(L, P, rs, C, fmt, fmt, rate, False, n_streams=S, device=local_rank, n_groups=1)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)          # torch events / copies and the engine share one stream
    eng.set_stream(stream.cuda_stream)
    Ct = S * C
    assert eng.set_coeff(make_filters(Ct, L * P, first=rank * Ct), P) == 0
    ring = 4
    host_in = [torch.from_numpy(noise_block(1000 * rank + b, S, L, C)).contiguous().pin_memory() for b in range(ring)]
    dev_in = [h.cuda(non_blocking=True) for h in host_in]
    dev_out = torch.empty(S * L * C, dtype=torch.float64, device="cuda")
    host_out = torch.empty(S * L * C, dtype=torch.float64).pin_memory()
    torch.cuda.synchronize()

    # prefill the delay line so that every timed step convolves all P partitions, then W warm-up steps
    for b in range(P):
        eng.run_device(dev_in[b % ring], dev_out)
    warm_out2 = torch.empty_like(dev_out)
    for b in range(0, max(W, 2), 2):       # warm-up through the two-block entry point (its buffers are allocated here)
        eng.run_device_pair(dev_in[b % ring], dev_in[(b + 1) % ring], dev_out, warm_out2)
    warm_outs = [dev_out, warm_out2, torch.empty_like(dev_out), torch.empty_like(dev_out)]
    eng.run_device_quad(dev_in, warm_outs)  # and the four-block one
    assert eng.sync() == 0
    del warm_outs

    # ---- device-resident throughput: EXACTLY K steps between barrier+sync, CUDA events on the launch stream.
    # A throughput caller has the next block at hand, so the steps go through the two-block entry point: both forward
    # transforms, ONE partition-sum launch that reads every coefficient spectrum once for both blocks, both inverse
    # transforms. "staged": the engine's stage pipeline (the pair sum stays on the engine's stream, where its CUDA-event
    # time is taken; the transforms of the next / previous pair run beside it on two side streams). "serial": the same
    # kernels back to back on one stream (each kernel's event time is its own: step shares). "single": one block per call.
    dev_out2 = torch.empty_like(dev_out)
    dev_outs = [dev_out, dev_out2] + [torch.empty_like(dev_out) for _ in range(2)]
    dev_outs8 = dev_outs + [torch.empty_like(dev_out) for _ in range(4)]

    def device_pass(mode):
        """K steps: 'staged_oct' / 'serial_oct' eight blocks per call, 'staged_quad' / 'serial_quad' four, 'staged' / 'serial'
        two, 'single' one; what does not fill a call goes through the next smaller one."""
        per = 8 if mode.endswith("oct") else (4 if mode.endswith("quad") else (1 if mode == "single" else 2))
        staged = mode.startswith("staged")
        eng.set_profiling(max(K // per, 1) + 2)
        n0 = pkg.kernel_launch_count()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        b = 0
        if per == 8:
            for b in range(0, K - 7, 8):
                eng.run_device_oct(dev_in + dev_in, dev_outs8, staged=staged)
            b = K - K % 8
        if per >= 4:
            for b in range(b, K - 3, 4):
                eng.run_device_quad(dev_in, dev_outs, staged=staged)
            b = K - K % 4
        if per >= 2:
            for b2 in range(b, K - 1, 2):
                eng.run_device_pair(dev_in[b2 % ring], dev_in[(b2 + 1) % ring], dev_out, dev_out2, pipelined=("staged" if staged else False))
            if K % 2:
                eng.run_device(dev_in[(K - 1) % ring], dev_out)
            eng.join()                      # the engine's stream waits for the side streams (no host wait)
        else:
            for b in range(K):
                eng.run_device(dev_in[b % ring], dev_out)
        e1.record(stream)
        assert eng.sync() == 0
        barrier()
        n = pkg.kernel_launch_count() - n0
        ms = max_over_ranks(e0.elapsed_time(e1))
        pr, npr = eng.get_profile()
        return ms, n, pr, npr

    def output_check():
        """The timed passes never look at their output: afterwards, run the SAME eight blocks through the stage pipeline
        (one eight-block call) on one engine and block by block on a second engine in the same state, and compare."""
        chk = pkg.Brutefir(L, P, rs, C, fmt, fmt, rate, False, n_streams=S, device=local_rank, n_groups=1)
        chk.set_stream(stream.cuda_stream)
        assert chk.set_coeff(make_filters(Ct, L * P, first=rank * Ct), P) == 0
        e2 = pkg.Brutefir(L, P, rs, C, fmt, fmt, rate, False, n_streams=S, device=local_rank, n_groups=1)
        e2.set_stream(stream.cuda_stream)
        assert e2.set_coeff(make_filters(Ct, L * P, first=rank * Ct), P) == 0
        o = [torch.empty_like(dev_out) for _ in range(16)]
        for b in range(P + 2):
            chk.run_device(dev_in[b % ring], o[0])
            e2.run_device(dev_in[b % ring], o[8])
        torch.cuda.synchronize()
        b = P + 2
        for k in range(8):
            chk.run_device(dev_in[(b + k) % ring], o[k])
        e2.run_device_oct([dev_in[(b + k) % ring] for k in range(8)], o[8:16], staged=True)
        e2.join()
        assert chk.sync() == 0 and e2.sync() == 0
        errs = []
        for k in range(8):
            d = (o[8 + k] - o[k]).double()
            errs.append(float(torch.sqrt(torch.mean(d * d) / torch.mean(o[k].double() ** 2))))
        chk.close()
        e2.close()
        return {"rel_rms_staged_oct_vs_single_blocks": errs, "ok": max(errs) < 1e-12,
                "note": "same eight blocks through bfir_run_device_oct (staged) and through eight bfir_run_device calls"}

    device_pass("staged_oct")               # warm-up of the stage pipeline (allocates its accumulators)
    for nb in (1, 2, 4, 8):
        eng.get_mac_profile(nb)
    sampler.busy.set()
    ms_total, launches, _, _ = device_pass("staged_oct")
    sampler.busy.clear()
    value = n_gpus * Ct * L * K / (ms_total * 1e-3) / 1e6
    mac8_sum, mac8_n = eng.get_mac_profile(8)            # the eight-block launches inside the timed region of `value`
    mac4_in_value = eng.get_mac_profile(4)               # (K % 8 >= 4: one four-block launch rides along)
    out_check = output_check()
    ms_serial8, _, _, _ = device_pass("serial_oct")
    mac8s_sum, mac8s_n = eng.get_mac_profile(8)
    device_pass("staged_quad")
    ms_quad_staged, _, prof_staged, nprof_staged = device_pass("staged_quad")
    ms_serial, _, prof, nprof = device_pass("serial_quad")
    # two blocks per call (staged and back to back), and one block per call (what a real-time caller gets; the
    # per-block partition sum of SURVEY 8d)
    device_pass("staged")
    ms_pair_staged, _, prof_pair_staged, nprof_pair_staged = device_pass("staged")
    ms_pair_serial, _, prof_pair, nprof_pair = device_pass("serial")
    ms_single, _, prof_single, nprof_single = device_pass("single")
    mac_split = eng.get_mac_split()
    quad_split = eng.get_quad_split()

    # the same device-resident work the way the end-to-end path runs it: 8 stream groups, no join between calls
    # (a group that is done with its blocks starts the next ones while others still convolve, so transforms run under
    # the partition sums of other groups).
    eng.set_groups(min(8, S))
    for b in range(0, 4, 2):
        eng.run_device_pair(dev_in[b % ring], dev_in[(b + 1) % ring], dev_out, dev_out2, pipelined=True)
    assert eng.sync() == 0
    barrier()
    ev0g, ev1g = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0g.record(stream)
    for b in range(0, K - 1, 2):
        eng.run_device_pair(dev_in[b % ring], dev_in[(b + 1) % ring], dev_out, dev_out2, pipelined=True)
    eng.join()
    ev1g.record(stream)
    assert eng.sync() == 0
    barrier()
    ms_grouped = max_over_ranks(ev0g.elapsed_time(ev1g))
    kp = K - K % 2
    value_grouped = {"value": n_gpus * Ct * L * kp / (ms_grouped * 1e-3) / 1e6, "ms_per_step": ms_grouped / kp,
                     "stream_groups": eng.get_groups(), "api": "bfir_run_device_pair(pipelined) + bfir_join"}

    # ---- end to end on pinned host buffers, every step: H2D of the step's input block, the kernels, D2H of its
    # output block. e2e.value: bfir_run_async_pair / bfir_wait (two blocks per call, DEPTH calls in flight, stream
    # groups with their own copy streams), median of 3 passes of K steps (the pass is half host work, and the hosts
    # of this pool are noisy). Beside it: one block per call (bfir_run_async) and the reference's synchronous run().
    DEPTH = 3
    QDEPTH = 3      # four-block calls in flight before the host waits for the oldest (measured: 2 leaves the copy streams starving now and then)

    def e2e_pass(engine, ins, outs, steps, sync_groups, async_groups):
        engine.set_groups(min(sync_groups, S))
        for b in range(3):
            rc, _ = engine.run(ins[b % len(ins)], outs[0])
            assert rc == 0
        barrier()
        t0 = time.perf_counter()
        for b in range(steps):
            rc, _ = engine.run(ins[b % len(ins)], outs[0])
        torch.cuda.synchronize()
        t_sync = time.perf_counter() - t0
        assert rc == 0
        engine.set_groups(min(async_groups, S))
        groups = engine.get_groups()
        ni, no = len(ins), len(outs)

        def async_pass(pairs):
            barrier()
            tickets = []
            t0 = time.perf_counter()
            if pairs:
                for k in range(steps // 2):
                    b = 2 * k
                    tickets.append(engine.run_async_pair(ins[b % ni], ins[(b + 1) % ni], outs[b % no], outs[(b + 1) % no]))
                    if k >= DEPTH:
                        assert engine.wait(tickets[k - DEPTH]) == 0   # blocks 2(k-DEPTH), +1 are now in their host buffers
                if steps % 2:
                    tickets.append(engine.run_async(ins[(steps - 1) % ni], outs[(steps - 1) % no]))
            else:
                for b in range(steps):
                    tickets.append(engine.run_async(ins[b % ni], outs[b % no]))
                    if b >= 2 * DEPTH:
                        assert engine.wait(tickets[b - 2 * DEPTH]) == 0
            assert engine.wait(tickets[-1]) == 0
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            barrier()
            return max_over_ranks(dt)
        async_pass(True)                                             # warm-up (allocates the staging ring)
        t_pairs = sorted(async_pass(True) for _ in range(3))
        t_single = async_pass(False)

        # four blocks per call through the stage pipeline (bfir_run_async_quad): ONE stream group -- whole-block copies
        # alternating between two copy streams each way, seven streams in all --, QDEPTH + 1 calls in flight
        def quad_pass():
            barrier()
            tickets = []
            t0 = time.perf_counter()
            for k in range(steps // 4):
                b = 4 * k
                tickets.append(engine.run_async_quad([ins[(b + j) % ni] for j in range(4)], [outs[(b + j) % no] for j in range(4)]))
                if k >= QDEPTH:
                    assert engine.wait(tickets[k - QDEPTH]) == 0
            for b in range(steps - steps % 4, steps):
                tickets.append(engine.run_async(ins[b % ni], outs[b % no]))
            assert engine.wait(tickets[-1]) == 0
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            barrier()
            return max_over_ranks(dt)
        engine.set_groups(1)
        quad_pass()
        t_quads = sorted(quad_pass() for _ in range(3))
        return max_over_ranks(t_sync), t_pairs, t_single, groups, t_quads

    e2e_steps = max(K, 400) - max(K, 400) % 4   # fill and drain of the copy pipeline are a fixed cost: time at least 400 steps
    n_host = max(2 * (DEPTH + 1), 4 * (QDEPTH + 1))
    host_ins = host_in + [torch.from_numpy(noise_block(1000 * rank + 50 + b, S, L, C)).contiguous().pin_memory() for b in range(n_host - ring)]
    host_outs = [host_out] + [torch.empty(S * L * C, dtype=torch.float64).pin_memory() for _ in range(n_host - 1)]
    np_in, np_outs = [h.numpy() for h in host_ins], [h.numpy() for h in host_outs]
    sampler.busy.set()
    t_sync, t_pairs, t_single, e2e_groups, t_quads = e2e_pass(eng, np_in, np_outs, e2e_steps, 4, 4)
    sampler.busy.clear()
    e2e_quads = t_quads[1] <= t_pairs[1]     # the headline is the faster of the two pipelined host paths (medians of 3 passes)
    t_e2e = t_quads[1] if e2e_quads else t_pairs[1]
    e2e_value = n_gpus * Ct * L * e2e_steps / t_e2e / 1e6
    e2e_sync_value = n_gpus * Ct * L * e2e_steps / t_sync / 1e6
    e2e_single_value = n_gpus * Ct * L * e2e_steps / t_single / 1e6
    checksum = float(host_out.numpy()[:1024].sum())

    # the link alone, same buffers, same process: every step's 8 MiB in and 8 MiB out as plain copies on two streams with
    # no kernel between them -- the floor of any end-to-end step on this box (PCIe duplex rate with this buffer ring)
    def link_only(steps):
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        d_i = [torch.empty_like(dev_in[0]) for _ in range(2)]
        d_o = [torch.empty_like(dev_out) for _ in range(2)]
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for b in range(steps):
            with torch.cuda.stream(s_in):
                d_i[b & 1].copy_(host_ins[b % n_host], non_blocking=True)
            with torch.cuda.stream(s_out):
                host_outs[b % n_host].copy_(d_o[b & 1], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        barrier()
        return max_over_ranks(dt)
    link_only(20)
    t_link = sorted(link_only(e2e_steps) for _ in range(3))[1]

    # ---- the same end-to-end step with the product's I/O format: foo_dsp_bfir constructs the engine with REALSIZE 8
    # and FLOAT_LE in/out (foo_dsp_bfir.cpp:279-286), which halves the PCIe bytes of the FLOAT64_LE headline run
    e2e_f32 = None
    if rank == 0 and not args.no_latency and world == 1:
        eng.close()
        e32 = pkg.Brutefir(L, P, rs, C, 8, 8, rate, False, n_streams=S, device=local_rank, n_groups=min(4, S))
        e32.set_stream(stream.cuda_stream)
        assert e32.set_coeff(make_filters(Ct, L * P, first=rank * Ct), P) == 0
        n32 = [h.float().pin_memory().numpy() for h in host_ins]
        no32 = [torch.empty(S * L * C, dtype=torch.float32).pin_memory().numpy() for _ in range(n_host)]
        for b in range(P):
            e32.run(n32[b % ring], no32[0])
        ts32, tp32, t1_32, _, tq32 = e2e_pass(e32, n32, no32, e2e_steps, 4, 4)
        tb32 = min(tp32[1], tq32[1])
        e2e_f32 = {"value": Ct * L * e2e_steps / tb32 / 1e6, "unit": "Msamples/s (this rank only)", "ms_per_step": 1e3 * tb32 / e2e_steps,
                   "two_blocks_per_call_ms_per_step": 1e3 * tp32[1] / e2e_steps, "four_blocks_per_call_ms_per_step": 1e3 * tq32[1] / e2e_steps,
                   "one_block_per_call": {"value": Ct * L * e2e_steps / t1_32 / 1e6, "ms_per_step": 1e3 * t1_32 / e2e_steps},
                   "sync_run": {"value": Ct * L * e2e_steps / ts32 / 1e6, "ms_per_step": 1e3 * ts32 / e2e_steps},
                   "h2d_bytes_per_step": S * L * C * 4, "d2h_bytes_per_step": S * L * C * 4,
                   "note": "FLOAT_LE in/out around the double-precision engine, as the plug-in runs it"}
        e32.close()

    # ---- roofline of the dominant kernel (partition MAC)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    # Bytes per launch. SURVEY 8d counts one channel-block of the partition sum as B_mac = (2P+1) N rs: P coefficient
    # spectra + P delay-line spectra in, one accumulated spectrum out. A launch of the FOUR-BLOCK kernel convolves four
    # consecutive blocks of every channel; all four use the same P coefficient spectra and, of the delay line, the
    # P + 3 spectra X[t+3] .. X[t-P+1] between them, so the bytes that MUST move are (P + (P+3) + 4) N rs = (2P+7) N rs
    # per channel -- that is what `achieved` / `frac` are taken on (roofline fraction <= ~1). The kernel moves
    # (2P + 3 SPLIT + 4) N rs (each of the SPLIT partition runs re-reads three boundary spectra; SPLIT = 1 here), ncu's
    # DRAM bytes per launch are `traffic`. 4 x B_mac, what four one-block launches (and the reference's access
    # pattern) would move, is kept under `vs_reference_access_pattern`.
    b_mac = (2 * P + 1) * (2 * L) * rs * Ct
    nquads = max(nprof, 1)
    mac_ms = mac8_sum / max(mac8_n, 1)                      # one launch = EIGHT blocks of every channel; timed region of `value`
    mac8_serial_ms = mac8s_sum / max(mac8s_n, 1)
    must_move = (2 * P + 15) * (2 * L) * rs * Ct
    achieved = must_move / (mac_ms * 1e-3) / 1e9 if mac_ms > 0 else 0.0
    traffic, traffic_src = None, None       # DRAM read+write bytes per launch from the committed ncu --set full capture
    tpath = os.path.join(ROOT, "profiles", "mac_traffic.json")
    if os.path.exists(tpath) and S == 16:
        tj = json.load(open(tpath))
        if tj.get("kernel", "").startswith("partition_mac_oct_kernel<double"):
            traffic, traffic_src = tj["traffic_bytes_per_launch"], tj["source"]
    mac4_ms = prof_staged["mac_ms"] / max(nprof_staged, 1)
    mac4_serial_ms = prof["mac_ms"] / nquads
    must_move4 = (2 * P + 7) * (2 * L) * rs * Ct
    mac1_ms = prof_single["mac_ms"] / max(nprof_single, 1)
    mac2_ms = prof_pair_staged["mac_ms"] / max(nprof_pair_staged, 1)
    mac2_serial_ms = prof_pair["mac_ms"] / max(nprof_pair, 1)
    must_move2 = (2 * P + 3) * (2 * L) * rs * Ct

    def gbs(nbytes, ms):
        return nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src,
                "kernel": "partition_mac_oct_kernel<double,W=4,128>", "peak_source": peak_src,
                "algorithmic_bytes_per_launch": must_move, "avg_launch_ms": mac_ms, "launches_timed": mac8_n,
                "note": "one launch = the partition sums of EIGHT consecutive blocks of %d channels; algorithmic bytes = what must move for that: "
                        "(2P+15) N realsize per channel (P coefficient + P+7 delay-line spectra in, 8 accumulated spectra out). Timed with CUDA events on "
                        "the engine's stream inside the timed region of `value`, i.e. WITH the transforms of the neighbouring calls running beside it"
                        "%s" % (Ct, "; the %d steps that do not fill an eight-block call went through one four-block launch (%.4f ms)" % (K % 8, mac4_in_value[0] / max(mac4_in_value[1], 1)) if mac4_in_value[1] else ""),
                "vs_reference_access_pattern": {"bytes_per_launch": 8 * b_mac, "frac": gbs(8 * b_mac, mac_ms) / peak,
                                                "note": "8 x SURVEY 8d's B_mac = (2P+1) N realsize per channel-block: what eight one-block partition sums move; "
                                                        "above 1 because the eight-block kernel reads every coefficient spectrum once for all eight blocks"},
                "serial_pass": {"note": "the same eight-block calls joined after every call (nothing beside the sum)", "avg_launch_ms": mac8_serial_ms,
                                "frac": gbs(must_move, mac8_serial_ms) / peak,
                                "value": n_gpus * Ct * L * K / (ms_serial8 * 1e-3) / 1e6, "ms_per_step": ms_serial8 / K},
                "four_blocks_per_launch": {"kernel": "partition_mac_multi_kernel<double,NB=4,SPLIT=%d>" % quad_split,
                                           "algorithmic_bytes_per_launch": must_move4, "note": "(2P+7) N realsize per channel for four blocks",
                                           "staged": {"avg_launch_ms": mac4_ms, "frac": gbs(must_move4, mac4_ms) / peak,
                                                      "value": n_gpus * Ct * L * K / (ms_quad_staged * 1e-3) / 1e6, "ms_per_step": ms_quad_staged / K},
                                           "serial": {"avg_launch_ms": mac4_serial_ms, "frac": gbs(must_move4, mac4_serial_ms) / peak,
                                                      "value": n_gpus * Ct * L * K / (ms_serial * 1e-3) / 1e6, "ms_per_step": ms_serial / K,
                                                      "step_share": {k: v / nquads / 4 for k, v in prof.items()}}},
                "step_share": {k: v / nquads / 4 for k, v in prof.items()},
                "two_blocks_per_launch": {"kernel": "partition_mac_pair_kernel<double,SPLIT=%d,UNROLL=2>" % mac_split,
                                          "algorithmic_bytes_per_launch": must_move2, "note": "(2P+3) N realsize per channel for two blocks",
                                          "staged": {"avg_launch_ms": mac2_ms, "frac": gbs(must_move2, mac2_ms) / peak,
                                                     "value": n_gpus * Ct * L * K / (ms_pair_staged * 1e-3) / 1e6, "ms_per_step": ms_pair_staged / K},
                                          "serial": {"avg_launch_ms": mac2_serial_ms, "frac": gbs(must_move2, mac2_serial_ms) / peak,
                                                     "value": n_gpus * Ct * L * K / (ms_pair_serial * 1e-3) / 1e6, "ms_per_step": ms_pair_serial / K,
                                                     "step_share": {k: v / max(nprof_pair, 1) / 2 for k, v in prof_pair.items()}}},
                "one_block_per_launch": {"kernel": "partition_mac_kernel<double,SPLIT=%d,UNROLL=4>" % mac_split, "avg_launch_ms": mac1_ms,
                                         "algorithmic_bytes_per_launch": b_mac, "note": "SURVEY 8d's B_mac x channels: here it IS the minimum",
                                         "frac": gbs(b_mac, mac1_ms) / peak,
                                         "value": n_gpus * Ct * L * K / (ms_single * 1e-3) / 1e6, "ms_per_step": ms_single / K,
                                         "step_share": {k: v / max(nprof_single, 1) for k, v in prof_single.items()}}}

    # ---- single-stream block latency (p50/p99 of host-visible bfir_run), rank 0
    latency = None
    if rank == 0 and not args.no_latency:
        e1 = pkg.Brutefir(L, P, rs, C, fmt, fmt, rate, False, n_streams=1, device=local_rank)
        assert e1.set_coeff(make_filters(C, L * P), P) == 0
        hin = [np.ascontiguousarray(h.numpy()[0]) for h in host_in]
        pin = [torch.from_numpy(x).pin_memory() for x in hin]
        pout = torch.empty(L * C, dtype=torch.float64).pin_memory()
        for b in range(P + 20):
            e1.run(pin[b % ring].numpy(), pout.numpy())
        pin_np, pout_np = [x.numpy() for x in pin], pout.numpy()

        def timed_calls(n, gap):
            lat = []
            for b in range(n):
                if gap:                             # idle time a real-time host has between blocks (period 170.7 ms)
                    t1 = time.perf_counter() + gap
                    while time.perf_counter() < t1:
                        pass
                t0 = time.perf_counter()
                e1.run(pin_np[b % ring], pout_np)
                lat.append(time.perf_counter() - t0)
            return np.sort(np.array(lat)) * 1e3
        lat = timed_calls(10000, 300e-6)            # p99 over >= 10 000 timed bfir_run calls
        lat_b2b = timed_calls(3000, 0.0)            # no gap: the look-ahead partition sum is still on the critical path
        e1.set_profiling(1000)                      # device-only time of the three kernels, CUDA events
        for b in range(1000):
            e1.run(pin_np[b % ring], pout_np)
        dprof, dn = e1.get_profile()
        latency = {"streams": 1, "calls": len(lat), "p50_ms": float(lat[len(lat) // 2]), "p99_ms": float(lat[int(len(lat) * 0.99)]),
                   "max_ms": float(lat[-1]), "pacing_gap_ms": 0.3,
                   "back_to_back": {"calls": len(lat_b2b), "p50_ms": float(lat_b2b[len(lat_b2b) // 2]), "p99_ms": float(lat_b2b[int(len(lat_b2b) * 0.99)])},
                   "device_kernels_ms": sum(dprof.values()) / max(dn, 1),
                   "block_period_ms": 1e3 * L / rate}
        e1.close()

    sampler.stop_flag.set()
    sampler.join(timeout=2)

    # ---- the other BASELINE configurations (tools/bench_configs.py)
    configs, sharded = None, None
    if not args.no_configs:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_configs as bc
        sh = importlib.import_module("foo-dsp-bfir_b200.sharding")
        torch.cuda.set_stream(torch.cuda.default_stream())
        configs = {}
        if rank == 0 and world == 1:
            configs["cfg0"] = bc.latency_plain(pkg, torch, "cfg0: stereo 44.1 kHz float, 65536 taps, L 4096, P 16 (bfir_run, pinned host buffers)",
                                               4096, 16, 4, 2, 44100, calls=3000)
            configs["cfg0_s16_dither"] = bc.latency_plain(pkg, torch, "cfg0 with S16_LE output and dither on", 4096, 16, 4, 2, 44100, calls=1500,
                                                          out_fmt=pkg.S16_LE, dither=True)
            configs["cfg2"] = bc.latency_cfg2(pkg, torch, calls=1000)
            configs["dither_kernel"] = bc.dither_timing(pkg, torch, streams=1)
            configs["dither_kernel_64_streams"] = bc.dither_timing(pkg, torch, streams=64)
        configs["cfg3"] = bc.throughput_cfg3(pkg, torch, peak, total_streams=4096, world=world, rank=rank, steps=40,
                                             max_over_ranks=max_over_ranks, barrier=barrier)
        sharded = bc.partition_sharded(pkg, sh, torch, dist, rank, world, local_rank, blocks=20)
        torch.cuda.set_stream(stream)

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        r = cpu_reference(host_cores(), blocks=1, steps=8, warmup=1)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": n_gpus, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(S, n_gpus),
            "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": S * L * C * 8,
                    "d2h_bytes_per_step": S * L * C * 8, "ms_per_step": 1e3 * t_e2e / e2e_steps, "steps": e2e_steps,
                    "api": ("bfir_run_async_quad(4 pinned host in, 4 pinned host out) + bfir_wait: H2D + kernels + D2H of every block through the stage "
                            "pipeline (one stream group, two copy streams each way), %d calls in flight; median of 3 passes" % (QDEPTH + 1)) if e2e_quads else
                           ("bfir_run_async_pair(pinned host in x2, pinned host out x2) + bfir_wait: H2D + kernels + D2H of every block, "
                            "%d calls in flight, %d stream groups with their own copy streams; median of 3 passes" % (DEPTH, e2e_groups)),
                    "passes_ms_per_step": [1e3 * t / e2e_steps for t in (t_quads if e2e_quads else t_pairs)],
                    "four_blocks_per_call": {"value": n_gpus * Ct * L * e2e_steps / t_quads[1] / 1e6, "passes_ms_per_step": [1e3 * t / e2e_steps for t in t_quads],
                                             "api": "bfir_run_async_quad + bfir_wait, one stream group"},
                    "two_blocks_per_call": {"value": n_gpus * Ct * L * e2e_steps / t_pairs[1] / 1e6, "passes_ms_per_step": [1e3 * t / e2e_steps for t in t_pairs],
                                            "api": "bfir_run_async_pair + bfir_wait, %d stream groups" % e2e_groups},
                    "one_block_per_call": {"value": e2e_single_value, "ms_per_step": 1e3 * t_single / e2e_steps, "api": "bfir_run_async + bfir_wait"},
                    "sync_run": {"value": e2e_sync_value, "ms_per_step": 1e3 * t_sync / e2e_steps,
                                 "api": "bfir_run(host in, host out): the reference's synchronous run(), 4 stream groups"},
                    "checksum": checksum,
                    "link_only": {"ms_per_step": 1e3 * t_link / e2e_steps, "frac_of_link": t_link / t_e2e,
                                  "note": "the same host buffers moved H2D and D2H concurrently on two streams with no kernels (median of 3 passes): "
                                          "the floor of an end-to-end step on this box; frac_of_link = link-only time / e2e time"}},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "latency": latency,
            "e2e_product_io": e2e_f32, "value_pipelined": value_grouped,
            "output_check": out_check, "configs": configs, "partition_sharded": sharded,
            "clocks": sampler.summary(),
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
