#!/usr/bin/env python
"""Where does e2e lose time against the device-only loop? Variants of the frame loop, CUDA-event timed.
   python tools/e2e_diag.py            (also under torchrun for N > 1)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import torch.distributed as dist
import bench
from ingvio_b200.synth import WORKLOADS, SyntheticStream

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
torch.cuda.set_device(local); dev = torch.device(f"cuda:{local}")
wl = WORKLOADS["c2"]; B, K, W = 1184, 12, 3
st = SyntheticStream(wl, 32, seq0=rank * B)
prefill = wl.sw - 1
frames = [bench.frame_arrays(st.next_frame()) for _ in range(prefill + W + K)]
ts = torch.cuda.Stream(device=dev)
with torch.cuda.stream(ts):
    g = bench.make_filter(wl, B, st, ts, local)
    to_dev = lambda d: {k: torch.from_numpy(bench.tile_to(v, B)).to(dev) for k, v in d.items()}
    to_pin = lambda d: {k: torch.from_numpy(bench.tile_to(v, B)).pin_memory() for k, v in d.items()}
    for i in range(prefill):
        bench.run_step(g, to_dev(frames[i][0]), frames[i][1])
    g.synchronize()
    idx = list(range(prefill, prefill + W + K))
    devf = {i: to_dev(frames[i][0]) for i in idx}
    pinf = {i: to_pin(frames[i][0]) for i in idx}
    nx = g.state_size()
    xbuf = [torch.empty((B, nx), dtype=torch.float64).pin_memory() for _ in range(2)]
    tbuf = [torch.empty((B,), dtype=torch.float64).pin_memory() for _ in range(2)]

    def loop(table, readback):
        n = {"i": 0}
        def rb():
            i = n["i"]
            if readback == "pipelined":
                g.get_state_async(xbuf[i % 2]); g.cov_trace_async(tbuf[i % 2]); g.fence_record(i % 2)
                if i > 0: g.fence_wait((i - 1) % 2)
            elif readback == "sync":
                g.get_state(); g.cov_trace()
            n["i"] = i + 1
        for i in idx[:W]:
            bench.run_step(g, table[i], frames[i][1]); rb()
        torch.cuda.synchronize(dev)
        if world > 1: dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record(ts)
        for i in idx[W:]:
            bench.run_step(g, table[i], frames[i][1]); rb()
        if readback == "pipelined": g.fence_wait((n["i"] - 1) % 2)
        e1.record(ts); torch.cuda.synchronize(dev)
        wall = (time.perf_counter() - t0) / K * 1e3
        ms = e0.elapsed_time(e1) / K
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
        return ms, wall

    for name, table, rbk in (("device inputs, no read-back", devf, None), ("device inputs, pipelined read-back", devf, "pipelined"),
                             ("host inputs, no read-back", pinf, None), ("host inputs, pipelined read-back", pinf, "pipelined"),
                             ("host inputs, sync read-back", pinf, "sync")):
        ms, wall = loop(table, rbk)
        if rank == 0: print(f"{name:40s} {ms:7.3f} ms/step (events, max over ranks)  {wall:7.3f} ms/step wall rank0", flush=True)
g.close()
if world > 1:
    dist.barrier(); dist.destroy_process_group()
