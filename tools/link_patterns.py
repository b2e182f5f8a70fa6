#!/usr/bin/env python
"""Which part of the pipelined host path's structure costs link rate? 8 MiB H2D + 8 MiB D2H per step on two copy streams:
  A  two device buffers each way, no dependencies (the bench's link-only floor)
  B  rings of 12 device buffers (the engine's staging slots)
  C  B + a consumer kernel behind every H2D and a producer kernel in front of every D2H on a third stream, chained by events
     (what the engine's transforms do to the staging slots)
  D  C with the copies of four steps queued in bursts (the four-block calls)
    python tools/link_patterns.py"""
import json, os, sys, time
import torch
n = 16 * 8192 * 8
NH, STEPS = 12, 240
h_in = [torch.rand(n, dtype=torch.float64).pin_memory() for _ in range(NH)]
h_out = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(NH)]
s_in, s_out, s_k = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()


def run(ring, kernels, burst):
    d_in = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(ring)]
    d_out = [torch.rand(n, dtype=torch.float64, device="cuda") for _ in range(ring)]
    ev_in = [torch.cuda.Event() for _ in range(ring)]
    ev_out = [torch.cuda.Event() for _ in range(ring)]
    ev_free_in = [torch.cuda.Event() for _ in range(ring)]
    ev_free_out = [torch.cuda.Event() for _ in range(ring)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for b0 in range(0, STEPS, burst):
        for b in range(b0, b0 + burst):
            k = b % ring
            with torch.cuda.stream(s_in):
                if kernels and b >= ring:
                    s_in.wait_event(ev_free_in[k])
                d_in[k].copy_(h_in[b % NH], non_blocking=True)
                ev_in[k].record(s_in)
        if kernels:
            for b in range(b0, b0 + burst):
                k = b % ring
                with torch.cuda.stream(s_k):
                    s_k.wait_event(ev_in[k])
                    d_in[k].mul_(1.0001)                      # consumer of the input slot
                    ev_free_in[k].record(s_k)
                    if b >= ring:
                        s_k.wait_event(ev_free_out[k])
                    d_out[k].add_(1.0)                        # producer of the output slot
                    ev_out[k].record(s_k)
        for b in range(b0, b0 + burst):
            k = b % ring
            with torch.cuda.stream(s_out):
                if kernels:
                    s_out.wait_event(ev_out[k])
                h_out[b % NH].copy_(d_out[k], non_blocking=True)
                ev_free_out[k].record(s_out)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / STEPS * 1e3


res = {}
for name, args in (("A", (2, False, 1)), ("B", (12, False, 1)), ("C", (12, True, 1)), ("D", (12, True, 4)), ("A2", (2, False, 1))):
    run(*args)
    res[name] = sorted(run(*args) for _ in range(3))[1]
print(json.dumps(res))
