#!/usr/bin/env python
"""What the host link gives for the end-to-end step's copies (8 MiB in + 8 MiB out per cfg1 x 16-stream block):
H2D alone, D2H alone, both at once on two streams; pinned memory from cudaHostAlloc with default flags and with
cudaHostAllocWriteCombined for the input side.   python tools/pcie_probe.py [MiB] [host buffers each way] [events]
`events`: every copy waits on an event of the other stream and records one behind it, as the engine's pipelined
host paths do (the events are long complete: this measures what the bookkeeping costs the link, not a dependency)."""
import ctypes, glob, json, os, sys, time
import torch

torch.cuda.init()
cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))
rt = ctypes.CDLL(cands[0])
vp, sz = ctypes.c_void_p, ctypes.c_size_t
rt.cudaHostAlloc.argtypes = [ctypes.POINTER(vp), sz, ctypes.c_uint]
rt.cudaMalloc.argtypes = [ctypes.POINTER(vp), sz]
rt.cudaMemcpyAsync.argtypes = [vp, vp, sz, ctypes.c_int, vp]
rt.cudaStreamCreate.argtypes = [ctypes.POINTER(vp)]
rt.cudaStreamSynchronize.argtypes = [vp]
rt.cudaEventCreateWithFlags.argtypes = [ctypes.POINTER(vp), ctypes.c_uint]
rt.cudaEventRecord.argtypes = [vp, vp]
rt.cudaStreamWaitEvent.argtypes = [vp, vp, ctypes.c_uint]
EVENTS = len(sys.argv) > 3 and sys.argv[3] == "events"
H2D, D2H = 1, 2
n = (int(sys.argv[1]) if len(sys.argv) > 1 else 8) << 20


def chk(e):
    assert e == 0, e


def host(flags):
    p = vp()
    chk(rt.cudaHostAlloc(ctypes.byref(p), n, flags))
    ctypes.memset(p, 1, n)
    return p


def dev():
    p = vp()
    chk(rt.cudaMalloc(ctypes.byref(p), n))
    return p


s1, s2 = vp(), vp()
chk(rt.cudaStreamCreate(ctypes.byref(s1))); chk(rt.cudaStreamCreate(ctypes.byref(s2)))
RING = int(sys.argv[2]) if len(sys.argv) > 2 else 1          # distinct host buffers cycled through, each way
d_in, d_out = dev(), dev()
evs = []
for _ in range(4):
    e = vp(); chk(rt.cudaEventCreateWithFlags(ctypes.byref(e), 2)); evs.append(e)
h_outs = [host(0) for _ in range(RING)]
for name, flags in (("default", 0), ("write_combined", 4)):
    h_ins = [host(flags) for _ in range(RING)]

    def run(h2d, d2h, iters=200):
        for it in range(iters + 20):
            if it == 20:
                rt.cudaStreamSynchronize(s1); rt.cudaStreamSynchronize(s2)
                t0 = time.perf_counter()
            if h2d:
                if EVENTS and it: chk(rt.cudaStreamWaitEvent(s1, evs[2 + (it - 1) % 2], 0))
                chk(rt.cudaMemcpyAsync(d_in, h_ins[it % RING], n, H2D, s1))
                if EVENTS: chk(rt.cudaEventRecord(evs[it % 2], s1))
            if d2h:
                if EVENTS and it and h2d: chk(rt.cudaStreamWaitEvent(s2, evs[(it - 1) % 2], 0))
                chk(rt.cudaMemcpyAsync(h_outs[it % RING], d_out, n, D2H, s2))
                if EVENTS: chk(rt.cudaEventRecord(evs[2 + it % 2], s2))
        rt.cudaStreamSynchronize(s1); rt.cudaStreamSynchronize(s2)
        return (time.perf_counter() - t0) / iters
    a, b, c = run(1, 0), run(0, 1), run(1, 1)
    print(json.dumps({"input_memory": name, "host_buffers_each_way": RING, "MiB_each_way": n >> 20, "events": EVENTS, "h2d_ms": a * 1e3, "d2h_ms": b * 1e3, "both_ms": c * 1e3,
                      "h2d_GBs": n / a / 1e9, "d2h_GBs": n / b / 1e9, "both_GBs_total": 2 * n / c / 1e9}))
