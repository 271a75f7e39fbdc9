#!/usr/bin/env python
"""bench.py -- EKF updates/sec (propagate + MSCKF visual + GNSS) on B200, BASELINE.json's metric.

One "step" = one frame cycle (10 IMU propagation steps, clone augmentation, MSCKF visual update over
F tracks seen in all SW clones, marginalisation of the oldest clone, GNSS pseudo-range/Doppler update;
call order of IngvioFilter::callbackMonoFrame, /root/reference/ingvio_estimator/src/IngvioFilter.cpp:143-231)
for EVERY one of the B independent sequences resident on each GPU. value = N_gpus * B * K / time.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--workload c2]
    python bench.py --impl reference ...       # CPU port of the reference path on all host cores

Timing: CUDA events on the handle's stream, barrier + synchronize on both sides, max over ranks.
`value`: inputs resident in HBM (device-pointer mode of the C-ABI). `e2e`: the same frames from pinned
HOST buffers through the same C-ABI (host-pointer mode: H2D copies inside the calls) plus a D2H read of
the per-sequence mean and trace(P) every step.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

from ingvio_b200.synth import WORKLOADS, SyntheticStream, R_C2I, P_C2I, R_CL2CR, P_CL2CR

METRIC = "ekf_updates_per_sec"
UNIT = "updates/s"
VISUAL_NOISE = 0.12
NOISE = dict(noise_g=0.004, noise_a=0.08, noise_bg=0.0002, noise_ba=0.008, noise_clockbias=0.2, noise_cb_rw=0.2)
GNSS_INIT = ((0, 1.0, 4.0), (1, -2.0, 4.0), (2, 0.5, 4.0), (3, 3.0, 4.0), (4, 0.1, 1.0), (5, 0.3, 0.015 ** 2))
COV_DIAG21 = np.array([0.0] * 3 + [0.0] * 3 + [0.25] * 3 + [0.01] * 3 + [0.01] * 3 + [1.8e-2] * 3 + [2e-3] * 3) ** 2


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=1184, help="independent sequences per GPU (148 SMs x 8)")
    ap.add_argument("--distinct", type=int, default=32, help="distinct synthetic streams per GPU, tiled to --batch")
    ap.add_argument("--cpu-sample-frames", type=int, default=200)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="fp64", choices=["fp64", "fp32_stack", "tf32_gram"],
                    help="fp32_stack: projected per-track blocks stored in single precision (IGV_PREC_FP32_STACK); tf32_gram: "
                         "the same, and their Gram matrix on the tcgen05 tensor cores, 3 x TF32 split (IGV_PREC_TF32_GRAM)")
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--no-c4", action="store_true", help="skip the c4 leg (64 sequences sharded over the ranks, per-frame gather)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# frames -> flat arrays the C-ABI consumes
# ------------------------------------------------------------------------------------------------
def frame_arrays(fr):
    g = fr.gnss
    d = dict(gyro=fr.gyro, accel=fr.accel, dt=fr.dt, pf=fr.pf_w, anchor=fr.anchor_slot.astype(np.int32), obs=fr.obs,
             mask=fr.obs_mask.astype(np.uint8), dof=(fr.obs_total.astype(np.int32) - 1))
    if g is not None:
        d.update(unit=g.unit, res_pos=g.res_pos, res_vel=g.res_vel, sig_psr=g.sigma_psr(), sig_dopp=g.sigma_dopp(),
                 sys=g.sys.astype(np.int32), Renu=g.R_enu2ecef.reshape(-1, 9))
    meta = dict(visual=fr.visual_mode is not None, marg=list(fr.marg_slots), max_valid=fr.max_valid, gnss=g is not None)
    return d, meta


def tile_to(a, B):
    reps = (B + a.shape[0] - 1) // a.shape[0]
    return np.ascontiguousarray(np.concatenate([a] * reps, 0)[:B])


def run_step(g, t, meta):
    """One frame cycle through the public host API (ingvio_b200.filter.BatchFilter -> C-ABI)."""
    from ingvio_b200 import capi
    g.propagate_imu(t["gyro"], t["accel"], t["dt"])
    g.augment_sliding_window_pose()
    if meta["visual"]:
        g.msckf_update(capi.VIS_ALL_OBS, t["pf"], t["anchor"], t["obs"], t["mask"], t["dof"], VISUAL_NOISE,
                       meta["max_valid"])
    for s in sorted(meta["marg"], reverse=True):
        g.marg_sliding_window_pose(s)
    if meta["gnss"]:
        g.gnss_update(t["unit"], t["res_pos"], t["res_vel"], t["sig_psr"], t["sig_dopp"], t["sys"], t["Renu"], 0, 0, 1)


class FixedFrame:
    """Device-resident argument buffers at FIXED addresses -- views into ONE packed allocation, so that a frame's inputs
    arrive with a single copy -- and the frame handed to the library as ONE call (igv_frame_step): from the third frame
    of a steady window the kernel sequence is a CUDA-graph replay."""

    def __init__(self, torch, dev, example):
        self.torch = torch
        self.layout, off = {}, 0
        for k, v in example.items():
            nbytes = v.numel() * v.element_size()
            self.layout[k] = (off, nbytes, v.dtype, tuple(v.shape))
            off += (nbytes + 255) & ~255
        self.nbytes = off
        self.store = torch.empty(off, dtype=torch.uint8, device=dev)
        self.buf = {k: self.store[o:o + n].view(dt).view(sh) for k, (o, n, dt, sh) in self.layout.items()}

    def pack(self, src, pinned):
        """One frame's inputs in the packed layout (host-pinned or on the device of `src`)."""
        torch = self.torch
        first = next(iter(src.values()))
        out = torch.empty(self.nbytes, dtype=torch.uint8, device="cpu" if pinned else first.device)
        if pinned:
            out = out.pin_memory()
        for k, (o, n, dt, sh) in self.layout.items():
            out[o:o + n].view(dt).view(sh).copy_(src[k])
        return out

    def load(self, packed):
        self.store.copy_(packed, non_blocking=True)      # ONE copy per frame

    def step(self, g, meta):
        from ingvio_b200 import capi
        b = self.buf
        vis = None
        if meta["visual"]:
            vis = dict(mode=capi.VIS_ALL_OBS, pf_w=b["pf"], anchor_slot=b["anchor"], obs=b["obs"], obs_mask=b["mask"],
                       chi2_dof=b["dof"], noise=VISUAL_NOISE, max_valid=meta["max_valid"])
        gn = None
        if meta["gnss"]:
            gn = dict(unit=b["unit"], res_pos=b["res_pos"], res_vel=b["res_vel"], sigma_psr=b["sig_psr"], sigma_dopp=b["sig_dopp"],
                      sys=b["sys"], R_enu2ecef=b["Renu"], is_adjust_yof=0, chi2_test=0, strong_reject=1)
        g.frame_step(b["gyro"], b["accel"], b["dt"], visual=vis, marg_slots=sorted(meta["marg"], reverse=True), gnss=gn)


def make_filter(wl, B, stream_obj, torch_stream, device):
    from ingvio_b200.filter import BatchFilter
    g = BatchFilter(B, wl.sw, max(wl.feats, 1), max(wl.sats, 1), stereo=wl.stereo, device=device,
                    stream=torch_stream.cuda_stream, noise=NOISE, T_cl2cr=(R_CL2CR, P_CL2CR), chi2_max_dof=160)
    ini = stream_obj.initial_state()
    t = lambda a: tile_to(a, B)
    g.init_state_and_cov(t(ini["R"].reshape(-1, 9)), t(ini["p"]), t(ini["v"]), t(ini["bg"]), t(ini["ba"]),
                         np.tile(R_C2I.reshape(1, 9), (B, 1)), np.tile(P_C2I, (B, 1)), COV_DIAG21)
    if wl.sats > 0:
        for gt, val, cov in GNSS_INIT:
            g.add_gnss_variable(gt, val, cov)
    return g


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("uuid,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, uuid):
        self.uuid = uuid
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        mine = [r for r in rows if len(r) >= 9 and (self.uuid is None or self.uuid in r[0])] or rows
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in mine:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
            except Exception:
                continue
            for k, nm in enumerate(names):
                if r[5 + k].strip().lower().startswith("active"):
                    reasons.add(nm)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=float(max(smax)) if smax else None,
                    reasons=sorted(reasons), samples=len(sm))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle's C++ port of the reference path (the reference itself cannot be built here)
# ------------------------------------------------------------------------------------------------
def _port_filter(wl, st):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from cpu_port.port import CpuPortFilter
    from ingvio_b200.filter import chi2_table
    f = CpuPortFilter([NOISE["noise_g"], NOISE["noise_a"], NOISE["noise_bg"], NOISE["noise_ba"], NOISE["noise_clockbias"],
                       NOISE["noise_cb_rw"]], [0, 0, -9.8], (R_CL2CR, P_CL2CR), wl.stereo, chi2_table(160, 0.95))
    ini = st.initial_state()
    f.init(ini["R"][0], ini["p"][0], ini["v"][0], ini["bg"][0], ini["ba"][0], R_C2I, P_C2I, COV_DIAG21)
    if wl.sats > 0:
        for gt, val, cov in GNSS_INIT:
            f.add_gnss(gt, val, cov)
    return f


def cpu_single_thread(wl, n_frames, seq0=0):
    """updates/s of ONE sequence on ONE host core (the reference is single-threaded, IngvioNode.cpp:36)."""
    st = SyntheticStream(wl, 1, seq0=seq0)
    f = _port_filter(wl, st)
    for _ in range(wl.sw - 1 + 2):
        f.step(st.next_frame().seq(0), VISUAL_NOISE)
    frames = [st.next_frame().seq(0) for _ in range(n_frames)]
    t0 = time.perf_counter()
    for fr in frames:
        f.step(fr, VISUAL_NOISE)
    dt = time.perf_counter() - t0
    return n_frames / dt, dt


def reference_build_rate():
    """Frames/s of the REFERENCE's own code (oracle/_ref/ref_driver: its unmodified translation units on the stand-in
    dense linear algebra of oracle/ref_shim) over the BASELINE-sized recorded stream of tests/ref_pin.py (window 11, 150
    tracks per image, visual path only). Reported as context beside the port, NOT a timing baseline: the stand-in Eigen is
    unoptimised, and the stream is the tracker-message one (no GNSS, real track lifetimes). None if the binary is absent."""
    drv = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    if not os.path.exists(drv):
        return None
    try:
        for p_ in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
            if p_ not in sys.path:
                sys.path.insert(0, p_)
        import ref_pin
        from test_cpp_updaters import _stream, _write_input
        big = dict(ref_pin.BIG, n_frames=60)
        wl, fp, st, frames = _stream(False, False, **big)
        with tempfile.TemporaryDirectory() as d:
            fin, fout = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
            _write_input(fin, wl, fp, st, frames, False)
            t0 = time.perf_counter()
            r = subprocess.run([drv, fin, fout], capture_output=True, text=True, timeout=300)
            dt = time.perf_counter() - t0
        if r.returncode != 0:
            return None
        return {"frames_per_sec": big["n_frames"] / dt, "frames": big["n_frames"], "seconds": dt,
                "what": "oracle/_ref/ref_driver: unmodified reference sources (propagate, augment, track table, triangulation, "
                        "RemoveLost + SwMarg updates, marginalise) on the stand-in linear algebra, window 11, 150 tracks/image, 1 core"}
    except Exception as e:
        return {"error": f"{type(e).__name__}: {e}"[:200]}


def usable_cores():
    """Host threads this process may really use: affinity mask capped by the cgroup CPU quota."""
    n = len(os.sched_getaffinity(0))
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(float(quota) / float(period))))
    except Exception:
        pass
    env = os.environ.get("IGV_REF_CORES")
    if env:
        n = min(n, int(env))
    return max(1, n)


def _ref_worker(args):
    wname, seq, warm, steps, barrier = args
    wl = WORKLOADS[wname]
    st = SyntheticStream(wl, 1, seq0=seq)
    f = _port_filter(wl, st)
    for _ in range(wl.sw - 1 + warm):
        f.step(st.next_frame().seq(0), VISUAL_NOISE)
    frames = [st.next_frame().seq(0) for _ in range(steps)]
    barrier.wait()
    t0 = time.perf_counter()
    for fr in frames:
        f.step(fr, VISUAL_NOISE)
    t1 = time.perf_counter()
    barrier.wait()
    return t0, t1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    wl = WORKLOADS[args.workload]
    cores = usable_cores()
    steps, warm = args.steps, max(args.warmup, 3)
    ctx = mp.get_context("fork")
    mgr = ctx.Manager()
    barrier = mgr.Barrier(cores)
    with ctx.Pool(cores) as pool:
        res = pool.map(_ref_worker, [(args.workload, i, warm, steps, barrier) for i in range(cores)], chunksize=1)
    t0 = min(r[0] for r in res)
    t1 = max(r[1] for r in res)
    wall = t1 - t0
    value = cores * steps / wall
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": wall / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{wl.name}: {'stereo' if wl.stereo else 'mono'} SW={wl.sw} F={wl.feats} S={wl.sats} "
                               f"(N={wl.dim}), one sequence per host core, {cores} sequences",
                   "step": "one frame cycle of every sequence"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{cores} sequences x {steps} frames after {wl.sw - 1 + warm} untimed frames; "
                                   "C++ port oracle/cpu_port (the reference needs Eigen/SuiteSparse/Boost/ROS, absent here)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def track_table_extra(wl, B, dev, ts, local, rank):
    """Extra (outside the metric): the device track table fed with tracker messages of wl.feats measurements
    per frame (15 % of the tracks replaced every frame, shuffled order), steady-state window. Times, per frame, the calls
    that replace the MapServer bookkeeping: collect, mark-lost, the two track selections (lost / seen-at), clean,
    re-anchor, erase-invalid. Triangulation and the updates themselves are measured elsewhere."""
    import torch
    from ingvio_b200 import capi
    from ingvio_b200.filter import BatchFilter
    M, SW, T = wl.feats, wl.sw, 4 * wl.feats
    rng = np.random.default_rng(99 + rank)
    g = BatchFilter(B, SW, max(M, 1), 1, stereo=wl.stereo, stream=ts.cuda_stream, device=local)
    try:
        eye = np.tile(np.eye(3).reshape(1, 9), (B, 1))
        z = np.zeros((B, 3))
        g.init_state_and_cov(eye, z, z, z, z, eye, z, np.full(21, 1e-2))
        g.create_map_server(T)
        n_frames = SW + 6
        cur = np.tile(np.arange(1, M + 1, dtype=np.uint64), (B, 1))
        nxt = M + 1
        msgs = []
        for _ in range(n_frames):
            repl = rng.random((B, M)) < 0.15
            fresh = (nxt + np.arange(M, dtype=np.uint64))[None, :].repeat(B, 0)
            nxt += M
            cur = np.where(repl, fresh, cur)
            perm = rng.permuted(np.tile(np.arange(M), (B, 1)), axis=1)
            ids = np.take_along_axis(cur, perm, 1)
            msgs.append((torch.full((B,), M, dtype=torch.int32, device=dev), torch.from_numpy(ids.astype(np.int64)).to(dev),
                         torch.from_numpy(rng.standard_normal((B, M, wl.rho)) * 0.3).to(dev)))
        F = max(M, 1)
        z_ = lambda shape, dt: torch.zeros(shape, dtype=dt, device=dev)   # noqa: E731
        out = dict(track_entry=z_((B, F), torch.int32), n_sel=z_((B,), torch.int32), track_id=z_((B, F), torch.int32),
                   obs=z_((B, F, SW, wl.rho), torch.float64), mask_all=z_((B, F, SW), torch.uint8),
                   mask_upd=z_((B, F, SW), torch.uint8), anchor_slot=z_((B, F), torch.int32),
                   chi2_dof=z_((B, F), torch.int32), feat_ok=z_((B, F), torch.uint8))
        Rd = torch.from_numpy(eye).to(dev)
        torch.cuda.synchronize(dev)
        ms_total, timed = 0.0, 0
        C_ = __import__("ctypes")
        prof_ms, prof_cnt = (C_.c_double * 8)(), (C_.c_longlong * 8)()
        lost_acc, seen_acc = z_((B,), torch.int32), z_((B,), torch.int32)   # same stream as the handle: ordered
        for k, (n_d, ids_d, uv_d) in enumerate(msgs):
            pd = torch.from_numpy(np.tile(np.array([0.3 * k, 0.0, 0.0]), (B, 1))).to(dev)
            torch.cuda.synchronize(dev)
            g.augment_sliding_window_pose_cov(Rd, Rd, pd)
            full = g.num_clones() >= SW
            if k == SW + 1:   # per-launch CUDA events of the library (family "other" = the track-table kernels)
                g.lib.igv_profile_enable(g.h, 1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ts)
            g.collect_meas(n_d, ids_d, uv_d)
            g.mark_marg_features()
            g.gather_tracks(capi.TRK_LOST, n_feats=F, obs_slots=SW, out=out)
            if full:
                lost_acc.add_(out["n_sel"])
            g.erase_tracks(out["track_entry"])
            if full:
                g.gather_tracks(capi.TRK_SEEN_AT, selected_slots=[0, SW // 2], n_feats=F, obs_slots=SW, out=out)
                seen_acc.add_(out["n_sel"])
                g.clean_obs_at([0])
                g.change_msckf_anchor([0], 0.0)
            g.erase_invalid_features(0.2)
            e1.record(ts)
            torch.cuda.synchronize(dev)
            if full:
                g.marg_sliding_window_pose(0)
                if k >= SW + 1:
                    ms_total += e0.elapsed_time(e1)
                    timed += 1
        ms = ms_total / max(1, timed)
        g.lib.igv_profile_read(g.h, prof_ms, prof_cnt, 1)
        g.lib.igv_profile_enable(g.h, 0)
        dev_ms, dev_launches = prof_ms[7] / max(1, timed), prof_cnt[7] / max(1, timed)
        nf = max(1, (n_frames - SW + 1) * B)
        lost, seen = int(lost_acc.sum().item()), int(seen_acc.sum().item())
        return {"ms_per_step": ms, "device_ms_per_step": dev_ms, "launches_per_step": dev_launches,
                "measurements_per_sec": B * M / (ms * 1e-3), "measurements_per_sec_device": B * M / (max(dev_ms, 1e-9) * 1e-3),
                "tracks_table_entries": T,
                "lost_per_frame": lost / nf, "seen_at_per_frame": seen / nf,
                "note": "igv_tracks_collect + mark_lost + gather(lost) + erase + gather(seen-at) + clean_obs + change_anchor + "
                        "erase_invalid per frame, device pointers; ms_per_step is bracketed by events around the Python calls (host submit "
                        "bound), device_ms_per_step sums the per-launch events of the track kernels (incl. the marginalisation "
                        "hook); outside the metric (SURVEY 8f-4)"}
    finally:
        g.close()


def c4_sharded(args, wl_name, world, rank, local, dev, ts, dist, torch, total=64):
    """BASELINE configs[3]: 64 independent c2 sequences block-sharded over the ranks (64 / N per GPU), frames from pinned
    HOST buffers through the C-ABI, and the per-frame NCCL all-gather of every sequence's read-out (R 9 + p 3 + trace(P) +
    flags = 14 doubles, 7 KB per frame at 64 sequences) INSIDE the timed region; rank 0 copies the gathered
    block to the host every frame. Latency-bound: the point of this leg is the small-batch regime, not throughput."""
    from ingvio_b200.sharding import gather_readouts_device, shard_range
    wl = WORKLOADS[wl_name]
    lo, hi = shard_range(total, world, rank)
    Bl = hi - lo
    K, W = max(args.steps, 20), max(args.warmup, 3)
    prefill = wl.sw - 1
    st = SyntheticStream(wl, Bl, seq0=100000 + lo)      # sequence b is the same stream whatever the sharding
    frames = [frame_arrays(st.next_frame()) for _ in range(prefill + W + K)]
    with torch.cuda.stream(ts):
        g = make_filter(wl, Bl, st, ts, local)
        pin = [({k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in f[0].items()}, f[1]) for f in frames]
        for a, m in pin[:prefill]:
            run_step(g, a, m)
        nx = g.state_size()
        xdev = torch.zeros((Bl, nx), dtype=torch.float64, device=dev)
        tdev = torch.zeros((Bl,), dtype=torch.float64, device=dev)
        host = torch.zeros((total, 14), dtype=torch.float64).pin_memory()
        checks = {"sum": 0.0}

        ff = FixedFrame(torch, dev, pin[prefill][0])
        packed = {i: ff.pack(pin[i][0], pinned=True) for i in range(prefill, len(pin))}

        def frame(a, m):
            ff.load(a)                              # ONE H2D copy of this frame's packed inputs from pinned memory
            ff.step(g, m)                           # ONE C-ABI call; a CUDA-graph replay in the steady state
            g.get_state_async(xdev)                 # the read-out stays in HBM
            g.cov_trace_async(tdev)
            ro = torch.cat([xdev[:, 0:12], tdev[:, None], torch.zeros((Bl, 1), dtype=torch.float64, device=dev)], 1)
            full = gather_readouts_device(ro, total, dist if world > 1 else None)
            if rank == 0:
                host.copy_(full, non_blocking=True)
                ts.synchronize()
                checks["sum"] += float(host[:, 12].sum())      # the host consumes the gathered read-out every frame

        for i in range(prefill, prefill + W):
            frame(packed[i], pin[i][1])
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = g.launch_count
        t0 = time.perf_counter()
        e0.record(ts)
        for i in range(prefill + W, len(pin)):
            frame(packed[i], pin[i][1])
        e1.record(ts)
        torch.cuda.synchronize(dev)
        wall_ms = (time.perf_counter() - t0) * 1e3
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        launches = g.launch_count - l0
        if world > 1:
            t = torch.tensor([ms, wall_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall_ms = float(t[0].item()), float(t[1].item())
        tr = g.cov_trace()
        assert np.all(np.isfinite(tr)) and np.all(tr > 0), "c4: filter diverged"
        h2d = sum(v.numel() * v.element_size() for v in pin[prefill][0].values())
        g_replays = g.graph_replays
        g.close()
    return {"workload": f"c4: {total} independent c2 sequences (mono SW={wl.sw} F={wl.feats} S={wl.sats}) block-sharded over "
                        f"{world} GPU(s) = {Bl} per GPU; per-frame NCCL all-gather of pose + trace(P) + flags inside the timed region",
            "value": total * K / (ms * 1e-3), "unit": UNIT, "per_gpu": total * K / (ms * 1e-3) / world,
            "ms_per_frame": ms / K, "wall_ms_per_frame": wall_ms / K, "frames": K, "warmup": W,
            "sequences_per_gpu": Bl, "gather": "all_gather_into_tensor, 14 doubles per sequence, every frame" if world > 1 else
            "single rank: no collective", "h2d_bytes_per_frame": int(h2d), "d2h_bytes_per_frame": total * 14 * 8,
            "gpu_launches_per_frame": launches / K, "graph_replays": g_replays, "submit": "igv_frame_step: one call per frame, CUDA-graph replay",
            "timing": "CUDA events on the filter's stream, max over ranks"}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist
    from ingvio_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"   # keep stdout to the one JSON line (no "NCCL version" banner)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (the CUDA path is the product; no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    wl = WORKLOADS[args.workload]
    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    D = max(1, min(args.distinct, B))
    prefill = wl.sw - 1

    # ---- synthetic streams (distinct seeds per rank), tiled to B sequences ----
    st = SyntheticStream(wl, D, seq0=rank * B)
    n_frames = prefill + 3 * (W + K)
    frames = [frame_arrays(st.next_frame()) for _ in range(n_frames)]
    ts = torch.cuda.Stream(device=dev)
    tdt = {np.dtype("float64"): torch.float64, np.dtype("int32"): torch.int32, np.dtype("uint8"): torch.uint8}

    def to_dev(d):
        return {k: torch.from_numpy(tile_to(v, B)).to(dev, non_blocking=False) for k, v in d.items()}

    def to_pinned(d):
        return {k: torch.from_numpy(tile_to(v, B)).pin_memory() for k, v in d.items()}

    with torch.cuda.stream(ts):
        g = make_filter(wl, B, st, ts, local)
        if args.precision == "fp32_stack":
            g.set_precision(capi.PREC_FP32_STACK)
        elif args.precision == "tf32_gram":
            g.set_precision(capi.PREC_TF32_GRAM)
        # fill the sliding window (untimed)
        for i in range(prefill):
            run_step(g, to_dev(frames[i][0]), frames[i][1])
        g.synchronize()
        seg = lambda k: list(range(prefill + k * (W + K), prefill + (k + 1) * (W + K)))
        dev_frames = {i: to_dev(frames[i][0]) for i in seg(0) + seg(2)}
        pin_frames = {i: to_pinned(frames[i][0]) for i in seg(1)}
        h2d_bytes = sum(v.numel() * v.element_size() for v in pin_frames[seg(1)[0]].values())

        def barrier():
            torch.cuda.synchronize(dev)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)

        def timed(idx, table, after_step=None, before_stop=None):
            for i in idx[:W]:
                run_step(g, table[i], frames[i][1])
                if after_step:
                    after_step()
            if before_stop:
                before_stop()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0 = g.launch_count
            e0.record(ts)
            for i in idx[W:]:
                run_step(g, table[i], frames[i][1])
                if after_step:
                    after_step()
            if before_stop:
                before_stop()
            e1.record(ts)
            barrier()
            ms = e0.elapsed_time(e1)
            if world > 1:
                t = torch.tensor([ms], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            return ms, g.launch_count - l0

        uuid = None
        try:
            uuid = str(torch.cuda.get_device_properties(dev).uuid)
        except Exception:
            pass
        # ---- (1) value: inputs resident in HBM ----
        clk = ClockSampler(uuid)
        clk.start()
        ms_dev, launches = timed(seg(0), dev_frames)
        clocks = clk.stop()
        # ---- (2) e2e: pinned host inputs through the same API + D2H read of mean and trace(P) EVERY step ----
        # The read-back is pipelined by one frame (igv_*_async + fences): the host submits frame k+1 (its bulk
        # host->device copies travel on the library's copy stream) before it consumes the result of frame k.
        nx = g.state_size()
        xbuf = [torch.empty((B, nx), dtype=torch.float64).pin_memory() for _ in range(2)]
        tbuf = [torch.empty((B,), dtype=torch.float64).pin_memory() for _ in range(2)]
        d2h = {"n": xbuf[0].numel() * 8 + tbuf[0].numel() * 8, "i": 0, "sum": 0.0}

        host = {"submit_s": 0.0, "wait_s": 0.0, "n": 0}

        def consume(slot):
            t0 = time.perf_counter()
            g.fence_wait(slot)
            host["wait_s"] += time.perf_counter() - t0
            d2h["sum"] += float(tbuf[slot][0]) + float(xbuf[slot][0, 9])   # the host really reads the result

        def read_back():
            i = d2h["i"]
            g.get_state_async(xbuf[i % 2])
            g.cov_trace_async(tbuf[i % 2])
            g.fence_record(i % 2)
            if i > 0:
                consume((i - 1) % 2)
            d2h["i"] = i + 1

        def drain():
            if d2h["i"] > 0:
                consume((d2h["i"] - 1) % 2)
            d2h["i"] = 0

        ms_e2e, _ = timed(seg(1), pin_frames, after_step=read_back, before_stop=drain)
        # what bounds e2e: raw host->device bandwidth of the same pinned buffers, and the host's own submit time
        big = max(pin_frames[seg(1)[0]].values(), key=lambda v: v.numel() * v.element_size())
        dst = torch.empty_like(big, device=dev)
        dst.copy_(big, non_blocking=True)
        torch.cuda.synchronize(dev)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(ts)
        for _ in range(5):
            dst.copy_(big, non_blocking=True)
        c1.record(ts)
        torch.cuda.synchronize(dev)
        h2d_gbs = 5 * big.numel() * big.element_size() / (c0.elapsed_time(c1) * 1e-3) / 1e9
        del dst
        # ---- (3) per-kernel-family device times (CUDA events on the launching stream) ----
        lib = g.lib
        import ctypes as C
        for i in seg(2)[:W]:
            run_step(g, dev_frames[i], frames[i][1])
        g.synchronize()
        lib.igv_profile_enable(g.h, 1)
        acc_out = torch.zeros(B, dtype=torch.int32, device=dev)
        for i in seg(2)[W:]:
            run_step(g, dev_frames[i], frames[i][1])
        ms_arr = (C.c_double * 8)()
        cnt_arr = (C.c_longlong * 8)()
        lib.igv_profile_read(g.h, ms_arr, cnt_arr, 1)
        lib.igv_profile_enable(g.h, 0)
        fam = {n: dict(ms=ms_arr[k], launches=int(cnt_arr[k])) for k, n in enumerate(capi.KERNEL_FAMILIES)}
        # ---- (4) extra (not part of the metric): the "next" row, device-side triangulation of the same tracks ----
        fi = seg(2)[-1]
        pf_d = torch.zeros((B, wl.feats, 3), dtype=torch.float64, device=dev)
        ok_d = torch.zeros((B, wl.feats), dtype=torch.uint8, device=dev)
        tri_ms = None
        if wl.feats > 0 and g.num_clones() >= 5:
            for _ in range(2):
                g.triangulate(dev_frames[fi]["obs"], dev_frames[fi]["mask"], dev_frames[fi]["anchor"], pf_out=pf_d, ok_out=ok_d)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ts)
            for _ in range(5):
                g.triangulate(dev_frames[fi]["obs"], dev_frames[fi]["mask"], dev_frames[fi]["anchor"], pf_out=pf_d, ok_out=ok_d)
            e1.record(ts)
            torch.cuda.synchronize(dev)
            tri_ms = e0.elapsed_time(e1) / 5
            tri_ok = float(ok_d.float().mean().item())
        # ---- (5) extra: the GNSS front end on the device: ephemerides -> satellite states -> residuals ----
        gfe_ms = None
        if wl.sats > 0:
            from ingvio_b200.synth import enu2ecef_rotation, geo2ecef, random_ephemerides
            S = wl.sats
            rng = np.random.default_rng(7 + rank)
            eph, gsys, tob, gpsr = random_ephemerides(rng, 1, S)
            rep = lambda a_, dt=torch.float64: torch.from_numpy(np.ascontiguousarray(np.repeat(a_, B, 0))).to(dev, dtype=dt)
            d_eph, d_tob, d_psr, d_sys = rep(eph), rep(tob), rep(gpsr), rep(gsys, torch.int32)
            sat = {k: torch.zeros(sh, dtype=torch.float64, device=dev)
                   for k, sh in dict(sat_pos=(B, S, 3), sat_vel=(B, S, 3), sat_clk=(B, S, 3), ttx_rel=(B, S)).items()}
            res = {k: torch.zeros(sh, dtype=torch.float64, device=dev)
                   for k, sh in dict(unit=(B, S, 3), res_pos=(B, S), res_vel=(B, S), sigma_psr=(B, S), sigma_dopp=(B, S),
                                     azel=(B, S, 2), atmos=(B, S, 2)).items()}
            d_obs = torch.stack([d_psr, torch.zeros_like(d_psr), torch.full_like(d_psr, 1575.42e6)], -1).contiguous()
            d_std = torch.ones((B, S, 3), dtype=torch.float64, device=dev)
            d_ttx = torch.full((B, S, 2), 100.0, dtype=torch.float64, device=dev)
            T12 = np.concatenate([enu2ecef_rotation(22.3, 114.2).reshape(9), geo2ecef(22.3, 114.2, 40.0)])
            d_T = torch.from_numpy(np.tile(T12, (B, 1))).to(dev)
            d_ion = torch.from_numpy(np.tile(np.array([0.1118e-7, -0.7451e-8, -0.5961e-7, 0.1192e-6, 0.1167e6, -0.2294e6,
                                                       -0.1311e6, 0.1049e7]), (B, 1))).to(dev)

            def gfe():
                g.sat_states(d_eph, d_tob, d_psr, d_sys, out=sat)
                g.gnss_residuals(sat["sat_pos"], sat["sat_vel"], sat["sat_clk"], d_obs, d_std, d_ttx, d_sys, d_T, d_ion, out=res)

            for _ in range(2):
                gfe()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ts)
            for _ in range(5):
                gfe()
            e1.record(ts)
            torch.cuda.synchronize(dev)
            gfe_ms = e0.elapsed_time(e1) / 5
        # ---- (6) extra: SLAM landmarks in the state (mono): delayed initialisation and the per-frame update ----
        lm_extra = None
        if wl.feats >= 8 and not wl.stereo and not args.no_latency:
            try:
                from ingvio_b200.filter import BatchFilter
                L = 8
                glm = BatchFilter(B, wl.sw, max(wl.feats, 1), max(wl.sats, 1), stereo=False, device=local, stream=ts.cuda_stream,
                                  noise=NOISE, T_cl2cr=(R_CL2CR, P_CL2CR), chi2_max_dof=160, max_landmarks=L)
                ini = st.initial_state()
                glm.init_state_and_cov(tile_to(ini["R"].reshape(-1, 9), B), tile_to(ini["p"], B), tile_to(ini["v"], B),
                                       tile_to(ini["bg"], B), tile_to(ini["ba"], B), np.tile(R_C2I.reshape(1, 9), (B, 1)),
                                       np.tile(P_C2I, (B, 1)), COV_DIAG21)
                if wl.sats > 0:
                    for gt, val, cov in GNSS_INIT:
                        glm.add_gnss_variable(gt, val, cov)
                for i in range(prefill):
                    run_step(glm, to_dev(frames[i][0]), frames[i][1])
                fr = to_dev(frames[prefill][0])      # the next frame: propagate + clone, then its tracks become landmarks
                glm.propagate_imu(fr["gyro"], fr["accel"], fr["dt"])
                glm.augment_sliding_window_pose()
                ncl = glm.num_clones()
                glm.synchronize()
                e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                for l in range(L):
                    if l == 1:
                        e0.record(ts)            # the first call grows the library's scratch arenas
                    glm.landmark_init(fr["pf"][:, l].contiguous(), ncl - 1, fr["obs"][:, l, :ncl, :2].contiguous(),
                                      fr["mask"][:, l, :ncl].contiguous(), VISUAL_NOISE)
                e1.record(ts)
                uv = fr["obs"][:, :L, ncl - 1, :2].contiguous()
                seen = torch.ones((B, L), dtype=torch.uint8, device=dev)
                for _ in range(5):
                    glm.landmark_update(uv, seen, VISUAL_NOISE)
                e2.record(ts)
                torch.cuda.synchronize(dev)
                lm_extra = {"landmarks_per_sequence": glm.num_landmarks(), "init_ms_per_landmark": e0.elapsed_time(e1) / (L - 1),
                            "update_ms": e1.elapsed_time(e2) / 5,
                            "landmark_updates_per_sec": B * glm.num_landmarks() / (e1.elapsed_time(e2) / 5 * 1e-3),
                            "note": "igv_landmark_init (delayed initialisation, k = 3) of 8 tracks and igv_landmark_update of the "
                                    "8 landmarks (2 x 24 rows each, chi^2 gate, one EKF update) on a second handle; outside the "
                                    "metric (SURVEY 8f-3)"}
                glm.close()
            except Exception as e:   # the extra must never take the bench line down
                lm_extra = {"error": repr(e)[:200]}
        flags = g.flags()
        tr = g.cov_trace()
        assert np.all(np.isfinite(tr)) and np.all(tr > 0), "filter diverged"
        # host-side cost of submitting one frame through the Python/ctypes/C-ABI stack (LAST: it replays frames, so the
        # filter state no longer matches the synthetic streams afterwards)
        t0 = time.perf_counter()
        for i in seg(1)[W:]:
            run_step(g, pin_frames[i], frames[i][1])
        host_submit_ms = (time.perf_counter() - t0) / K * 1e3
        g.synchronize()
        n_flag = int(np.count_nonzero(flags & 3))

    total_updates = world * B * K
    value = total_updates / (ms_dev * 1e-3)
    e2e_value = total_updates / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel ----
    # SURVEY.md section 8(d) per-sequence figures of the stages a kernel covers, x B sequences per launch.
    n = 6 * wl.sw
    q = wl.rho * wl.sw - 3
    m = wl.feats * q
    r_keep = min(m, n)
    feat_bytes = 96.0 * wl.sw + 8.0 * wl.rho * wl.feats * wl.sw + 24.0 * wl.feats + 8.0 * m * (n + 1)
    feat_flops = wl.feats * sum(4.0 * (wl.rho * wl.sw - k) * (n + 3 - k) for k in range(3)) + wl.feats * wl.sw * (wl.rho / 2) * 198.0
    qr_bytes1 = 8.0 * m * (n + 1) + 8.0 * r_keep * (n + 1)
    qr_flops1 = (2.0 * m * n * n - 2.0 / 3.0 * n ** 3 + 4.0 * m * n) if m > n else 0.0
    fused = g.last_visual_path() == 2
    fam_launch = lambda k: max(1, fam[k]["launches"])
    if fused:
        # one kernel does the per-track Jacobian / null space / gate AND the compression's accumulation: it is
        # charged with both stages' algorithmic work; the n x n factorisation (k_gram_factor) is family "qr"
        dom, dom_kernel = "features", "k_msckf_features<FUSE> (per-track Jacobian + null space + gate + Gram accumulation, DMMA)"
        dom_bytes, dom_flops = B * (feat_bytes + qr_bytes1), B * (feat_flops + qr_flops1)
        compulsory = B * (96.0 * wl.sw + 8.0 * wl.rho * wl.feats * wl.sw + 24.0 * wl.feats + 8.0 * n * n + 8.0 * (n + 1) * (n + 2) / 2)
    elif fam["qr"]["ms"] / fam_launch("qr") >= fam["features"]["ms"] / fam_launch("features"):
        dom, dom_kernel = "qr", ("k_gram_stream/k_gram_accum + k_gram_factor" if g.last_visual_path() == 1 else "k_qr_* (Householder)")
        dom_bytes, dom_flops, compulsory = B * qr_bytes1, B * qr_flops1, B * qr_bytes1
    else:
        dom, dom_kernel = "features", "k_msckf_features"
        dom_bytes, dom_flops, compulsory = B * feat_bytes, B * feat_flops, B * feat_bytes
    dom_ms = fam[dom]["ms"] / fam_launch(dom)
    hbm_peak, peak_src = peaks()
    achieved_gbs = dom_bytes / (dom_ms * 1e-3) / 1e9
    fp64_peak = C.c_double(0.0)
    lib.igv_measure_fp64_peak(local, C.byref(fp64_peak))      # warm-up call
    total_prof_ms = sum(v["ms"] for v in fam.values())
    traffic = None
    try:  # DRAM bytes per launch of this kernel from the committed ncu capture, if it is the same configuration
        tj = json.load(open(os.path.join(ROOT, "profiles", "dominant_traffic.json")))
        if tj.get("workload") == wl.name and int(tj.get("batch", -1)) == B and tj.get("family") == dom and bool(tj.get("fused")) == fused:
            traffic = float(tj["dram_bytes_per_launch"])
    except Exception:
        pass
    fp64_clk = ClockSampler(uuid)
    fp64_clk.start()
    best = 0.0
    for _ in range(12):                                        # ~0.5 s of probe under the clock sampler: the peak's clock record
        lib.igv_measure_fp64_peak(local, C.byref(fp64_peak))
        best = max(best, fp64_peak.value)
    fp64_peak.value = best
    fp64_clocks = fp64_clk.stop()
    fp64_achieved = dom_flops / (dom_ms * 1e-3) / 1e12
    # PRIMARY roofline: the FP64 pipe. The dominant kernel moves ~1 % of what HBM could carry in its run time (see
    # roofline_hbm.traffic), so bandwidth does not bind it; arithmetic issue does.
    roofline = {"kernel": dom_kernel, "bound": "fp64_pipe", "achieved": fp64_achieved, "peak": fp64_peak.value,
                "unit": "TFLOP/s", "frac": fp64_achieved / fp64_peak.value if fp64_peak.value else None,
                "traffic": traffic,
                "peak_source": "igv_measure_fp64_peak: DFMA probe on this GPU in this run (DMMA.8x8x4 measures the same rate); "
                               "MEASURED_PEAKS.json holds no FP64 figure",
                "peak_clocks": fp64_clocks,
                "algorithmic_flops_per_launch": dom_flops, "avg_launch_ms": dom_ms,
                "share_of_step": fam[dom]["ms"] / total_prof_ms if total_prof_ms else None,
                "note": "FLOPs = SURVEY 8(d) figures of the stages this kernel covers (dense null-space projection + Householder "
                        "QR of the stack) x sequences per launch; the Gram formulation executes fewer. traffic = dram bytes per "
                        "launch from the committed ncu capture (profiles/dominant_traffic.json)"}
    roofline_hbm = {"kernel": dom_kernel, "bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                    "frac": achieved_gbs / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                    "achieved_dram_gbs": (traffic / (dom_ms * 1e-3) / 1e9) if traffic else None,
                    "algorithmic_bytes_per_launch": dom_bytes, "compulsory_bytes_per_launch": compulsory,
                    "avg_launch_ms": dom_ms,
                    "note": "secondary: `achieved` uses the ALGORITHMIC bytes of SURVEY 8(d) (they include the projected stack's HBM "
                            "round trip, which the fused kernel never materialises); achieved_dram_gbs is what really moves"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp64": "f64", "fp32_stack": "f64 (projected stack stored as f32)",
                  "tf32_gram": "f64 (projected stack stored as f32, its Gram matrix as 3 x TF32 on tcgen05 with FP64 sums)"}[args.precision],
        "data": "synthetic",
        "config": {"workload": f"{wl.name}: {'stereo' if wl.stereo else 'mono'} SW={wl.sw} F={wl.feats} S={wl.sats} "
                               f"(N={wl.dim}), B={B} independent sequences per GPU",
                   "batch_per_gpu": B, "distinct_streams_per_gpu": D, "imu_steps_per_frame": 10,
                   "step": "one frame cycle (propagate x10, augment, MSCKF update, marginalise, GNSS update) of every sequence",
                   "parallelism": f"sequences sharded over {world} GPU(s), no data-path collective",
                   "l2": "per-step working set (stacked Jacobians + covariances) >> 126 MB L2; no explicit flush needed"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h["n"]),
                "ms_per_step": ms_e2e / K, "h2d_gbs_measured": h2d_gbs, "host_submit_ms_per_step": host_submit_ms,
                "host_wait_ms_per_step": host["wait_s"] / max(1, W + K) * 1e3,
                "note": "pipelined by one frame: the result of frame k is read after frame k+1 is submitted; bulk H2D on the "
                        "library's copy stream"},
        "gpu_launches": int(launches),
        "roofline": roofline, "roofline_hbm": roofline_hbm,
        "kernel_ms_per_step": {k: v["ms"] / K for k, v in fam.items()},
        "flagged_sequences": n_flag,
    }
    if tri_ms is not None:
        line["triangulation_extra"] = {"ms_per_step": tri_ms, "tracks_per_sec": B * wl.feats / (tri_ms * 1e-3),
                                       "accepted_fraction": tri_ok,
                                       "note": "igv_triangulate on the same tracks; outside the metric (SURVEY 8f-1)"}

    if lm_extra is not None:
        line["landmark_extra"] = lm_extra
    if gfe_ms is not None:
        line["gnss_frontend_extra"] = {"ms_per_step": gfe_ms, "satellites_per_sec": B * wl.sats / (gfe_ms * 1e-3),
                                       "note": "igv_sat_states + igv_gnss_residuals on synthetic ephemerides; outside the "
                                               "metric (SURVEY 8f-2)"}

    if rank == 0 and world == 1 and not args.no_latency and wl.feats > 0:
        try:
            with torch.cuda.stream(ts):
                line["track_table_extra"] = track_table_extra(wl, B, dev, ts, local, rank)
        except Exception as e:   # an extra must never cost the bench line
            line["track_table_extra"] = {"error": f"{type(e).__name__}: {e}"[:300]}

    if rank == 0 and world == 1 and not args.no_latency:
        # single-sequence latency (BASELINE configs[1] as one filter): B = 1 handle, row-split QR
        with torch.cuda.stream(ts):
            st1 = SyntheticStream(wl, 1, seq0=777)
            g1 = make_filter(wl, 1, st1, ts, local)
            fr1 = [frame_arrays(st1.next_frame()) for _ in range(prefill + W + K)]
            t1 = [({k: torch.from_numpy(v).to(dev) for k, v in f[0].items()}, f[1]) for f in fr1]
            for a, mta in t1[:prefill]:
                run_step(g1, a, mta)
            ff1 = FixedFrame(torch, dev, t1[prefill][0])
            p1 = [ff1.pack(a, pinned=False) for a, _ in t1]
            for i in range(prefill, prefill + W):
                ff1.load(p1[i])
                ff1.step(g1, t1[i][1])
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t_host = time.perf_counter()
            e0.record(ts)
            for i in range(prefill + W, len(t1)):
                ff1.load(p1[i])            # ONE device -> device refill of the fixed argument buffers
                ff1.step(g1, t1[i][1])     # one igv_frame_step call: graph replay
            e1.record(ts)
            host_ms = (time.perf_counter() - t_host) * 1e3 / K
            torch.cuda.synchronize(dev)
            line["single_sequence"] = {"ms_per_update": e0.elapsed_time(e1) / K, "updates_per_sec": K / (e0.elapsed_time(e1) * 1e-3),
                                       "host_submit_ms_per_update": host_ms, "graph_replays": g1.graph_replays,
                                       "submit": "igv_frame_step (one call per frame, CUDA-graph replay)"}
            g1.close()

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, secs = cpu_single_thread(wl, args.cpu_sample_frames)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": f"1 sequence, {args.cpu_sample_frames} steady-state frames of the same workload "
                                          f"({secs:.1f} s) on one host core; C++ port oracle/cpu_port "
                                          "(reference needs Eigen/SuiteSparse/Boost/ROS, absent here)"}
        rb = reference_build_rate()
        if rb is not None:
            line["cpu_baseline"]["reference_build"] = rb
    g.close()
    if not args.no_c4 and wl.name == "c2":
        try:
            line["c4_sharded"] = c4_sharded(args, "c2", world, rank, local, dev, ts, dist, torch)
        except Exception as e:   # an extra must never cost the bench line (every rank takes the same path)
            line["c4_sharded"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
