#!/usr/bin/env python3
"""bench.py -- ORBSLAMM hot-path benchmark on B200 (contract: see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W [--workload frontend|ba] [--impl reference]

frontend (default, BASELINE.json metric "ORB extract+match fps @1241x376"):
    one step = one new 1241x376 frame for each of `--streams` independent synthetic camera streams:
    ORB extraction (pyramid, FAST, quad-tree, orientation, rBRIEF) + SearchByProjection(Cur, Last) against the
    stream's previous frame (+ PoseOptimization once enabled).  value = frames/s with the images already in HBM;
    e2e = the same through the host-buffer C-ABI calls (H2D of the images, D2H of keypoints/descriptors/matches).
ba (metric "LocalBA LM iters/s @500KF/50k pts"): see bench_ba() below.

--impl reference times the CPU oracle (cv2 primitives + restated reference code) on all host cores.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from orbslamm_b200 import synth  # noqa: E402

CAM = synth.KITTI
TH_PROJ = 15.0          # Tracking.cc:925-930 (mono)
TRAFFIC_JSON = "r2i_traffic.json"
N_POOL = 5              # distinct frames per stream (steps cycle over frame pairs 1..N_POOL-1)


def ncu_traffic(group, kernel, scale_to=None):
    """dram__bytes_read + dram__bytes_write of one launch from the committed ncu capture (profiles/r2h_traffic.json, written by
    tools/ncu_summary.py traffic from the --set full capture of the final kernels), scaled linearly to this run's frames per launch; None when
    the capture has no such kernel."""
    p = os.path.join(ROOT, "profiles", TRAFFIC_JSON)
    try:
        d = json.load(open(p))[group]
        v = d["kernels"][kernel][0]["dram_bytes"]
        if scale_to is not None:
            v = v * scale_to / d["frames_per_launch"]
        return int(v)
    except Exception:
        return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm = [int(r[0]) for r in self.rows if r and r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def frontend_workload(n_streams, rank):
    """Synthetic streams: images [N_POOL, B, H, W] u8 and per-frame known shifts."""
    w, h = CAM["w"], CAM["h"]
    imgs = np.empty((N_POOL, n_streams, h, w), np.uint8)
    shifts = np.zeros((N_POOL, n_streams, 2), np.int32)
    for b in range(n_streams):
        fr, sh = synth.stream(w, h, N_POOL, stream_id=rank * 1000 + b)
        for t in range(N_POOL):
            imgs[t, b] = fr[t]; shifts[t, b] = sh[t]
    return imgs, shifts


def predicted_pose():
    T = np.eye(4, dtype=np.float32)
    T[:3, 3] = [0.02, -0.01, 0.03]
    return T


def build_queries(feats, shifts_next, rng):
    """Last-frame 'map points' for one stream frame: keypoints lifted to depth U(5, 50) m so that they project onto
    their shifted position in the next frame under the next frame's pose (identity here)."""
    fx, fy, cx, cy = CAM["fx"], CAM["fy"], CAM["cx"], CAM["cy"]
    n = len(feats["x"])
    z = rng.uniform(5, 50, n)
    dx, dy = shifts_next
    Xw = np.stack([(feats["x"] + dx - cx) * z / fx, (feats["y"] + dy - cy) * z / fy, z], 1).astype(np.float32)
    return Xw


def bench_frontend(args, rank, world):
    import torch
    import orbslamm_b200 as ob
    from orbslamm_b200 import build as obuild
    obuild.build()
    dev = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(dev)
    B = args.streams
    w, h = CAM["w"], CAM["h"]
    imgs, shifts = frontend_workload(B, rank)
    ex = ob.ORBextractor(CAM["nfeatures"], 1.2, 8, 20, 7, device=dev)
    mt = ob.ORBmatcher(0.9, True, device=dev)
    mt.set_stream(ex.stream())
    po = ob.Optimizer(device=dev)
    po.set_stream(ex.stream())
    L = ob.load()
    slab = ex.max_keypoints(w, h)
    sf = ex.GetScaleFactors()
    K4 = np.array([CAM["fx"], CAM["fy"], CAM["cx"], CAM["cy"]], np.float32)
    bounds = np.array([0, 0, w, h], np.float32)

    # ---- untimed pre-pass: extract every pool frame once to build the per-frame map-point queries
    rng = np.random.default_rng(1234 + rank)
    pitch = (w + 63) // 64 * 64
    d_imgs = torch.zeros((N_POOL, B, h, pitch), dtype=torch.uint8, device="cuda")
    d_imgs[:, :, :, :w] = torch.from_numpy(imgs).cuda()
    torch.cuda.synchronize()
    q_Xw = np.zeros((N_POOL, B, slab, 3), np.float32); q_oct = np.zeros((N_POOL, B, slab), np.int32)
    q_ang = np.zeros((N_POOL, B, slab), np.float32); q_desc = np.zeros((N_POOL, B, slab, 32), np.uint8)
    q_valid = np.zeros((N_POOL, B, slab), np.uint8); q_cnt = np.zeros((N_POOL, B), np.int32)
    host_feats = []
    for t in range(N_POOL):
        ex.extract_device(d_imgs[t].data_ptr(), B, w, h, pitch, h * pitch)
        r = ex.download(B, slab)
        host_feats.append(r)
        for b in range(B):
            n = int(r["counts"][b]); q_cnt[t, b] = n
            nxt = shifts[t + 1, b] if t + 1 < N_POOL else (0, 0)
            feats = dict(x=r["xy"][b, :n, 0], y=r["xy"][b, :n, 1])
            q_Xw[t, b, :n] = build_queries(feats, nxt, rng)
            q_oct[t, b, :n] = r["octave"][b, :n]; q_ang[t, b, :n] = r["angle"][b, :n]; q_desc[t, b, :n] = r["desc"][b, :n]
            q_valid[t, b, :n] = 1
    to_dev = lambda a: torch.from_numpy(a).cuda()
    dq = dict(Xw=to_dev(q_Xw), oct=to_dev(q_oct), ang=to_dev(q_ang), desc=to_dev(q_desc), valid=to_dev(q_valid), cnt=to_dev(q_cnt))
    Tcw = np.tile(predicted_pose().reshape(1, 16), (B, 1))     # motion-model prediction: slightly off the true (identity) pose
    d_Tcw0 = to_dev(Tcw); d_Tcw = to_dev(Tcw.copy()); d_sf = to_dev(sf)
    ils = ex.GetInverseScaleSigmaSquares(); d_ils = to_dev(ils)
    d_fout = torch.zeros((B, slab), dtype=torch.uint8, device="cuda"); d_ninl = torch.zeros(B, dtype=torch.int32, device="cuda")
    d_qvalid = torch.zeros((B, slab), dtype=torch.uint8, device="cuda")
    d_quv = torch.zeros((B, slab, 2), dtype=torch.float32, device="cuda"); d_qrad = torch.zeros((B, slab), dtype=torch.float32, device="cuda")
    d_qmn = torch.zeros((B, slab), dtype=torch.int32, device="cuda"); d_qmx = torch.zeros((B, slab), dtype=torch.int32, device="cuda")
    d_fm = torch.zeros((B, slab), dtype=torch.int32, device="cuda"); d_nm = torch.zeros(B, dtype=torch.int32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.ExternalStream(ex.stream(), device=dev)
    torch.cuda.synchronize()
    vp = ctypes.c_void_p

    def step_device(t):
        """extract frame t of every stream, match against frame t-1's map points -- all on ex.stream(), no host sync"""
        ex.extract_device(d_imgs[t].data_ptr(), B, w, h, pitch, h * pitch)
        v = ex.device_view()
        with torch.cuda.stream(stream):
            d_qvalid.copy_(dq["valid"][t - 1], non_blocking=True)
            d_fm.fill_(-1)
            d_Tcw.copy_(d_Tcw0, non_blocking=True)
        ob._check(L.orbm_project_last_frame(mt.handle, B, vp(d_Tcw.data_ptr()), K4.ctypes.data, bounds.ctypes.data, vp(d_sf.data_ptr()), len(sf),
                                            vp(dq["Xw"][t - 1].data_ptr()), vp(dq["oct"][t - 1].data_ptr()), vp(dq["cnt"][t - 1].data_ptr()), slab,
                                            TH_PROJ, vp(d_qvalid.data_ptr()), vp(d_quv.data_ptr()), vp(d_qrad.data_ptr()), vp(d_qmn.data_ptr()),
                                            vp(d_qmx.data_ptr()), 1))
        ob._check(L.orbm_search_by_projection(mt.handle, B, bounds.ctypes.data, vp(v.kp_xy), vp(v.kp_octave), vp(v.kp_angle), vp(v.desc), vp(v.counts),
                                              slab, vp(d_qvalid.data_ptr()), vp(d_quv.data_ptr()), vp(d_qrad.data_ptr()), vp(d_qmn.data_ptr()),
                                              vp(d_qmx.data_ptr()), vp(dq["ang"][t - 1].data_ptr()), vp(dq["desc"][t - 1].data_ptr()),
                                              vp(dq["cnt"][t - 1].data_ptr()), slab, 100, 0.0, 1, vp(d_fm.data_ptr()), vp(d_nm.data_ptr()), 1))
        ob._check(L.orbo_pose_optimization_matched(po.handle, B, vp(d_Tcw.data_ptr()), K4.ctypes.data, vp(v.kp_xy), vp(v.kp_octave), vp(v.counts), slab,
                                                   vp(d_fm.data_ptr()), vp(dq["Xw"][t - 1].data_ptr()), vp(dq["cnt"][t - 1].data_ptr()), slab,
                                                   vp(d_ils.data_ptr()), len(ils), vp(d_fout.data_ptr()), vp(d_ninl.data_ptr()), None, 1))

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing.  The streams of this GPU are served by `--instances` independent extractor / matcher /
    # optimizer triples, each on its own CUDA stream (the library is re-entrant; the reference runs one System per robot in one
    # process): the latency-bound per-frame kernels of one instance (quad-tree, sequential-equivalent match resolution, pose LM)
    # overlap the throughput-bound kernels of the others.  One step is timed fork-join on a parent stream with CUDA events.
    nI = max(1, min(args.instances, B))

    class Inst:
        def __init__(self, b0, b1):
            self.b0, self.b1, self.n = b0, b1, b1 - b0
            n = self.n
            self.ex = ob.ORBextractor(CAM["nfeatures"], 1.2, 8, 20, 7, device=dev)
            self.mt = ob.ORBmatcher(0.9, True, device=dev); self.mt.set_stream(self.ex.stream())
            self.po = ob.Optimizer(device=dev); self.po.set_stream(self.ex.stream())
            self.stream = torch.cuda.ExternalStream(self.ex.stream(), device=dev)
            z = lambda shape, dt: torch.zeros(shape, dtype=dt, device="cuda")
            self.T0 = d_Tcw0[b0:b1]; self.T = z((n, 16), torch.float32)
            self.fout = z((n, slab), torch.uint8); self.ninl = z((n,), torch.int32); self.qvalid = z((n, slab), torch.uint8)
            self.quv = z((n, slab, 2), torch.float32); self.qrad = z((n, slab), torch.float32)
            self.qmn = z((n, slab), torch.int32); self.qmx = z((n, slab), torch.int32)
            self.fm = z((n, slab), torch.int32); self.nm = z((n,), torch.int32)
            self.done = torch.cuda.Event()

        def step(self, t):
            b0, b1, n = self.b0, self.b1, self.n
            ex_, mt_, po_ = self.ex, self.mt, self.po
            ex_.extract_device(d_imgs[t][b0:b1].data_ptr(), n, w, h, pitch, h * pitch)
            v = ex_.device_view()
            q = {k: dq[k][t - 1][b0:b1] for k in dq}
            with torch.cuda.stream(self.stream):
                self.qvalid.copy_(q["valid"], non_blocking=True)
                self.fm.fill_(-1)
                self.T.copy_(self.T0, non_blocking=True)
            D = lambda x: vp(x.data_ptr())
            ob._check(L.orbm_project_last_frame(mt_.handle, n, D(self.T), K4.ctypes.data, bounds.ctypes.data, D(d_sf), len(sf), D(q["Xw"]), D(q["oct"]),
                                                D(q["cnt"]), slab, TH_PROJ, D(self.qvalid), D(self.quv), D(self.qrad), D(self.qmn), D(self.qmx), 1))
            ob._check(L.orbm_search_by_projection(mt_.handle, n, bounds.ctypes.data, vp(v.kp_xy), vp(v.kp_octave), vp(v.kp_angle), vp(v.desc), vp(v.counts),
                                                  slab, D(self.qvalid), D(self.quv), D(self.qrad), D(self.qmn), D(self.qmx), D(q["ang"]), D(q["desc"]),
                                                  D(q["cnt"]), slab, 100, 0.0, 1, D(self.fm), D(self.nm), 1))
            ob._check(L.orbo_pose_optimization_matched(po_.handle, n, D(self.T), K4.ctypes.data, vp(v.kp_xy), vp(v.kp_octave), vp(v.counts), slab,
                                                       D(self.fm), D(q["Xw"]), D(q["cnt"]), slab, D(d_ils), len(ils), D(self.fout), D(self.ninl), None, 1))

        def launches(self):
            return self.ex.kernel_launches() + self.mt.kernel_launches() + self.po.kernel_launches()

    # Two ways to use the instances:
    #   pipeline (default): every instance serves ALL streams of this GPU on alternating steps -- while instance A matches and optimises the poses of
    #       step i (latency-bound kernels: one CTA per frame), instance B already extracts the frames of step i+1 (throughput-bound kernels).  That is the
    #       software pipeline a tracker runs (the extraction of the next frame does not depend on the pose of the current one); the K steps are timed as
    #       ONE region (device events on a parent stream around all of them).  No L2 flush inside the region: one step reads 60 MB of new images out of
    #       a 298 MB pool and writes / re-reads a 0.6 GB working set (pyramids, blurred levels) -- inputs larger than the 126 MB L2.
    #   split: the streams are split over the instances, every step is a fork-join timed on its own, L2 flushed between steps.
    pipelined = args.mode == "pipeline" and nI > 1
    if pipelined:
        insts = [Inst(0, B) for _ in range(nI)]
    else:
        insts = [Inst(i * B // nI, (i + 1) * B // nI) for i in range(nI)]
    parent = torch.cuda.Stream(device=dev)

    def step_all(t, e0, e1):
        e0.record(parent)
        for it in insts:
            it.stream.wait_event(e0)
            it.step(t)
            it.done.record(it.stream)
            parent.wait_event(it.done)
        e1.record(parent)

    def run_pipelined(n_steps, e0, e1):
        e0.record(parent)
        for it in insts:
            it.stream.wait_event(e0)
        for i in range(n_steps):
            insts[i % nI].step(1 + i % (N_POOL - 1))
        for it in insts:
            it.done.record(it.stream)
            parent.wait_event(it.done)
        e1.record(parent)

    dummy = (torch.cuda.Event(), torch.cuda.Event())
    if pipelined:
        run_pipelined(max(args.warmup, nI), *dummy)
    else:
        for i in range(args.warmup):
            step_all(1 + i % (N_POOL - 1), *dummy)
    barrier()
    l0 = sum(it.launches() for it in insts)
    sampler = ClockSampler(dev); sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall = time.time()
    if pipelined:
        flush.fill_(1)
        torch.cuda.synchronize()
        run_pipelined(args.steps, *evs[0])
        barrier()
        total_ms = evs[0][0].elapsed_time(evs[0][1])
    else:
        for i in range(args.steps):
            flush.fill_(i & 0xff)                       # L2 flush between timed iterations (untimed, torch's stream)
            torch.cuda.synchronize()
            step_all(1 + i % (N_POOL - 1), *evs[i])
        barrier()
        total_ms = sum(a.elapsed_time(b) for a, b in evs)
    wall = time.time() - t_wall
    launches = sum(it.launches() for it in insts) - l0
    last = insts[(args.steps - 1) % nI] if pipelined else None
    ninl = (last.ninl if pipelined else torch.cat([it.ninl for it in insts])).cpu().numpy()
    nmatch = (last.nm if pipelined else torch.cat([it.nm for it in insts])).cpu().numpy()
    if world > 1:
        tt = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        total_ms = float(tt.item())
    fps = world * B * args.steps / (total_ms * 1e-3)

    # ---- per-kernel pass (roofline): ONE instance over all B streams, every kernel bracketed by CUDA events on its launching
    # stream, L2 flushed between steps -- kernel durations here are not inflated by another instance sharing the SMs
    for i in range(2):
        step_device(1 + i % (N_POOL - 1))
    barrier()
    ex.set_profiling(True)
    prof_steps = min(args.steps, 8)
    pev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(prof_steps)]
    for i in range(prof_steps):
        flush.fill_(i & 0xff)
        torch.cuda.synchronize()
        pev[i][0].record(stream)
        step_device(1 + i % (N_POOL - 1))
        pev[i][1].record(stream)
    barrier()
    prof_ms = sum(a.elapsed_time(b) for a, b in pev)
    ktimes = ex.kernel_times()
    ex.set_profiling(False)

    # ---- end-to-end through the host-buffer C-ABI: orbf_track_frames calls (pinned host images and map points in; keypoints,
    # descriptors, matches, poses and outlier flags out; the stages in between stay on the device).  Like the reference's
    # multi-robot binary (one System per robot, each with its own ORBextractor, in one process), the streams of this GPU are
    # served by `--e2e-workers` independent front-end instances (own handles, own CUDA stream, own host thread): while one
    # instance uploads or downloads, the others compute.
    nW = max(1, min(args.e2e_workers, B))
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    P = lambda tns: vp(tns.data_ptr())
    A = lambda arr: arr.ctypes.data
    bounds_w = [(i * B // nW, (i + 1) * B // nW) for i in range(nW)]

    class Worker:
        def __init__(self, b0, b1):
            self.b0, self.b1, self.n = b0, b1, b1 - b0
            n = self.n
            self.ex = ob.ORBextractor(CAM["nfeatures"], 1.2, 8, 20, 7, device=dev)
            self.mt = ob.ORBmatcher(0.9, True, device=dev); self.po = ob.Optimizer(device=dev)
            self.fe = ob.FrontEnd(self.ex, self.mt, self.po, device=dev)
            self.imgs = pin(imgs[:, b0:b1])
            self.q = dict(Xw=pin(q_Xw[:, b0:b1]), oct=pin(q_oct[:, b0:b1]), ang=pin(q_ang[:, b0:b1]), desc=pin(q_desc[:, b0:b1]),
                          valid=pin(q_valid[:, b0:b1]), cnt=pin(q_cnt[:, b0:b1]))
            e = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
            self.o_xy = e((n, slab, 2), torch.float32); self.o_ang = e((n, slab), torch.float32); self.o_resp = e((n, slab), torch.float32)
            self.o_oct = e((n, slab), torch.int32); self.o_size = e((n, slab), torch.float32); self.o_desc = e((n, slab, 32), torch.uint8)
            self.o_cnt = torch.zeros(n, dtype=torch.int32).pin_memory(); self.o_fm = e((n, slab), torch.int32)
            self.o_nm = torch.zeros(n, dtype=torch.int32).pin_memory(); self.o_out = e((n, slab), torch.uint8)
            self.o_ninl = torch.zeros(n, dtype=torch.int32).pin_memory(); self.h_T = pin(Tcw[b0:b1].copy())
            self.T0 = torch.from_numpy(Tcw[b0:b1].copy())
            self.err = None

        def step(self, t):
            self.h_T.copy_(self.T0)
            q = self.q
            ob._check(L.orbf_track_frames(self.fe.handle, P(self.imgs[t]), self.n, w, h, w, w * h, A(K4), A(sf), A(ils), len(sf), P(q["Xw"][t - 1]),
                                          P(q["oct"][t - 1]), P(q["ang"][t - 1]), P(q["desc"][t - 1]), P(q["valid"][t - 1]), P(q["cnt"][t - 1]), slab,
                                          TH_PROJ, 100, 1, P(self.h_T), P(self.o_xy), P(self.o_ang), P(self.o_resp), P(self.o_oct), P(self.o_size),
                                          P(self.o_desc), slab, P(self.o_cnt), P(self.o_fm), P(self.o_nm), P(self.o_out), P(self.o_ninl)))

        def run(self, steps, gate):
            try:
                torch.cuda.set_device(dev)
                gate.wait()
                for i in range(steps):
                    self.step(1 + i % (N_POOL - 1))
            except Exception as ex_:          # surfaced by the main thread
                self.err = ex_

    workers = [Worker(b0, b1) for b0, b1 in bounds_w]

    def run_all(steps):
        gate = threading.Barrier(nW + 1)
        ths = [threading.Thread(target=wk.run, args=(steps, gate)) for wk in workers]
        for th_ in ths: th_.start()
        barrier()
        t0 = time.perf_counter()
        gate.wait()
        for th_ in ths: th_.join()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        for wk in workers:
            if wk.err is not None: raise wk.err
        return dt

    run_all(min(args.warmup, 3))
    e2e_steps = args.steps if args.steps <= 20 else 20 + (args.steps - 20) % (N_POOL - 1)   # ends on the same pool frame as the device loop
    e2e_s = run_all(e2e_steps)
    barrier()
    clocks = sampler.stop()                         # sampled over both timed regions (device-resident loop and e2e loop)
    if world > 1:
        tt = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e_fps = world * B * e2e_steps / e2e_s
    # ---- single-call latency ("latency per frame must be reported separately"): the reference serves 1 or 2 live camera streams
    # (mono_kitti.cc / mono_kitti_dif-Seq.cc), so one orbf_track_frames call on 1 and on 2 frames -- host buffers in, host results out,
    # the call returns when the results are on the host -- is timed on its own, median of 30 calls, nothing else running on the GPU
    latency = {}
    for nb in (1, 2):
        if nb > B:
            continue
        wk = Worker(0, nb)
        for i in range(5):
            wk.step(1 + i % (N_POOL - 1))
        ts = []
        for i in range(30):
            t0 = time.perf_counter()
            wk.step(1 + i % (N_POOL - 1))
            ts.append(time.perf_counter() - t0)
        latency[f"streams_{nb}"] = round(float(np.median(ts)) * 1e3, 4)
        del wk
    latency["note"] = "median wall ms of one orbf_track_frames call (extract + match + pose) with pinned host buffers, copies included, 30 calls"
    o_cnt = torch.cat([wk.o_cnt for wk in workers]); o_nm = torch.cat([wk.o_nm for wk in workers]); o_ninl = torch.cat([wk.o_ninl for wk in workers])
    launches_e2e = sum(wk.ex.kernel_launches() + wk.mt.kernel_launches() + wk.po.kernel_launches() for wk in workers)
    nkp = int(np.mean(o_cnt.numpy()))
    assert np.array_equal(o_nm.numpy(), nmatch) and np.array_equal(o_ninl.numpy(), ninl), "host-buffer path and device-resident path disagree"
    se = B * slab                                   # slab entries per step
    h2d = B * w * h + se * (12 + 4 + 4 + 32 + 1) + B * 4 + B * 64 + 64     # images, last-frame map points, counts, poses, tables
    d2h = se * (8 + 4 * 4 + 32) + B * 4 + se * (4 + 1) + B * 8 + B * 64     # keypoints + descriptors (strided slabs), matches, flags, poses

    # ---- roofline of the dominant kernel
    hbm, how = peaks()
    lw = ctypes.c_int(); lh = ctypes.c_int(); px = []
    for l in range(8):
        L.orbx_level_size(ex.handle, w, h, l, ctypes.byref(lw), ctypes.byref(lh)); px.append(lw.value * lh.value)
    sum_px = sum(px)
    ncand = sum(len(ex.candidates(0, l)) for l in range(8))
    alg = {"resize_level": (sum_px - px[-1]) + (sum_px - px[0]),        # read levels 0..6, write levels 1..7 (all 7 launches)
           "fast_cells": sum_px + 8 * ncand, "blur7": 2 * sum_px, "octree": 8 * ncand + 8 * nkp,
           "orient_describe": nkp * (709 + 512 + 60)}
    dom = max(ktimes, key=lambda k: ktimes[k][0])
    dom_ms, dom_n = ktimes[dom]
    per_launch_ms = dom_ms / max(dom_n, 1) * (7 if dom == "resize_level" else 1)
    achieved = alg[dom] * B / (per_launch_ms * 1e-3) / 1e9
    roofline = {"kernel": dom, "bound": "hbm", "achieved": round(achieved, 2), "peak": hbm, "unit": "GB/s", "frac": round(achieved / hbm, 5),
                "traffic": ncu_traffic("frontend", dom, B) if w == 1241 else None, "traffic_source": f"profiles/{TRAFFIC_JSON} (ncu --set full, 128 frames per launch, scaled to this batch)",
                "peak_source": how, "algorithmic_bytes_per_launch": int(alg[dom] * B),
                "note": "k_fast_cells is bound by instruction issue / the integer ALU pipe (ncu: issue ~73 %, alu pipe ~77 % of peak, DRAM ~3 %), not by HBM; frac is reported against the HBM roofline as the contract asks",
                "avg_launch_ms": round(per_launch_ms, 4),
                "kernel_share_of_step": {k: round(v[0] / max(prof_ms, 1e-9), 4) for k, v in ktimes.items()},
                "measured_in": f"single-instance pass over all {B} streams ({prof_steps} steps, {round(prof_ms / prof_steps, 4)} ms/step), CUDA events around every launch; "
                               "blur7 runs on a side stream beside octree, so their two timed regions overlap (shares can add up to more than the step)"}

    out = {"metric": f"ORB extract+match fps @{w}x{h}", "value": round(fps, 1), "unit": "frames/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": round(total_ms / args.steps, 4), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "u8", "data": "synthetic",
           "config": frontend_config(args, world),
           "stats": {"keypoints_per_frame": nkp, "matches_per_frame": float(np.mean(nmatch)), "pose_inliers_per_frame": float(np.mean(ninl))},
           "e2e": {"value": round(e2e_fps, 1), "unit": "frames/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                   "workers": nW, "gpu_launches": int(launches_e2e),
                   "note": "orbf_track_frames on pinned host buffers; the GPU's streams are split over `workers` independent front-end instances "
                           "(own handles / CUDA stream / host thread), as the reference runs one System per robot in one process"},
           "latency_ms": latency,
           "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "wall_s": round(wall, 3)}
    return out


def frontend_config(args, world):
    """The workload description both arms print verbatim (`--impl reference` runs on THIS config); measured per-frame counts live in `stats`."""
    w, h = CAM["w"], CAM["h"]
    B = args.streams
    pitch = (w + 63) // 64 * 64
    nI = max(1, min(args.instances, B))
    pipelined = args.mode == "pipeline" and nI > 1
    return {"workload": f"{'KITTI' if w == 1241 else 'TUM'}-shape {w}x{h} synthetic streams, nFeatures={CAM['nfeatures']}: extract + SearchByProjection(Cur,Last) + PoseOptimization per frame",
            "streams_per_gpu": B, "frames_per_step": B * world, "instances_per_gpu": nI,
            "instances": ("software pipeline: every instance serves all streams on alternating steps (extract of step i+1 overlaps match + pose of step i); the K steps are one timed region"
                          if pipelined else "streams split over the instances, fork-join per step"),
            "l2": ("inputs larger than L2: each step reads %.0f MB of new images from a %.0f MB pool and a %.2f GB working set; no flush inside the timed region" % (B * h * pitch / 1e6, N_POOL * B * h * pitch / 1e6, 2 * B * 1.444e6 / 1e9 + B * h * pitch / 1e9))
                  if pipelined else "256 MiB flush buffer written between timed steps (untimed)",
            "parallelism": f"streams x{world}"}


# ------------------------------------------------------------------------------------------------
def cpu_frontend_frames(stream_id, n_frames):
    """Reference-faithful CPU front end for one stream (cv2 primitives + restated reference code), 1 thread."""
    import cv2
    import oracle
    from oracle import orb_cv2
    cv2.setNumThreads(1)
    frames, shifts = synth.stream(CAM["w"], CAM["h"], n_frames + 1, stream_id=stream_id)
    P = oracle.orb_params(CAM["nfeatures"], 1.2, 8, 20, 7)
    g = oracle.grid_params(0, 0, CAM["w"], CAM["h"])
    sf = np.array(list(P.scale)[:8], np.float32)
    K4 = np.array([CAM["fx"], CAM["fy"], CAM["cx"], CAM["cy"]], np.float32)
    rng = np.random.default_rng(stream_id)
    ils = np.array(list(P.inv_sigma2)[:8], np.float32)
    T_pred = predicted_pose()
    last = orb_cv2.extract(P, frames[0])
    t0 = time.perf_counter()
    for t in range(1, n_frames + 1):
        cur = orb_cv2.extract(P, frames[t])
        Xw = build_queries(last, shifts[t], rng)
        qv, uv, rad, mn, mx = oracle.project_last_frame(T_pred, K4, g, sf, Xw, last["octave"], TH_PROJ,
                                                         np.ones(len(Xw), np.uint8))
        fxy = np.stack([cur["x"], cur["y"]], 1)
        _, fm = oracle.search_by_projection(g, fxy, cur["octave"], cur["angle"], cur["desc"], qv, uv, rad, mn, mx,
                                            last["angle"], last["desc"], 100, 0.0, True)
        m = fm >= 0
        oracle.pose_optimization(T_pred, Xw[fm[m]], fxy[m], ils[cur["octave"][m]], K4)
        last = cur
    return n_frames, time.perf_counter() - t0


def _cpu_worker(a):
    return cpu_frontend_frames(*a)


def cpu_baseline_frontend(cores, frames_per_core):
    import multiprocessing as mp
    if cores == 1:
        n, s = cpu_frontend_frames(900, frames_per_core)
        return n / s, s
    with mp.get_context("fork").Pool(cores) as pool:
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, [(900 + i, frames_per_core) for i in range(cores)])
        wall = time.perf_counter() - t0
    # throughput of the parallel phase: every worker times its own loop; use the slowest
    n = sum(r[0] for r in res); s = max(r[1] for r in res)
    return n / s, wall


_REAL_STDOUT = None


def cpu_model():
    """CPU model and logical core count of the box the baseline ran on (SURVEY.md 8d asks for both next to the CPU number)."""
    name = "unknown"
    try:
        for l in open("/proc/cpuinfo"):
            if l.lower().startswith("model name"):
                name = l.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    return f"{name} ({os.cpu_count()} logical cores)"


def _own_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on communicator creation), so file
    descriptor 1 is pointed at stderr for the whole run and the result line goes to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


def main():
    _own_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--streams", type=int, default=128, help="independent camera streams per GPU (frames per step per GPU)")
    ap.add_argument("--instances", type=int, default=2, help="independent extractor/matcher/optimizer triples (own CUDA stream) per GPU in the device-resident leg")
    ap.add_argument("--mode", default="pipeline", choices=["pipeline", "split"], help="how the device-resident leg uses --instances (see bench_frontend)")
    ap.add_argument("--e2e-workers", type=int, default=4, help="independent front-end instances serving the streams of one GPU in the e2e leg")
    ap.add_argument("--workload", default="all", choices=["all", "frontend", "ba", "merge"],
                    help="all (default): the front-end line with the LocalBA leg as its \"ba\" sub-record -- both halves of BASELINE.json's metric")
    ap.add_argument("--camera", default="kitti", choices=["kitti", "tum"], help="kitti: 1241x376 / 2000 features (BASELINE.json metric); tum: 640x480 / 1000 features (configs[1])")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-frames", type=int, default=100, help="bounded CPU-baseline sample (frames)")
    ap.add_argument("--ba-kf", type=int, default=500)
    ap.add_argument("--ba-pts", type=int, default=50000)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    global CAM
    CAM = synth.TUM if args.camera == "tum" else synth.KITTI
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))

    if args.impl == "reference":
        if rank != 0:
            return
        if args.workload == "ba":
            from bench_ba import reference_line
            emit(reference_line(args))
            return
        if args.workload == "merge":
            from bench_merge import reference_merge
            emit(reference_merge(args))
            return
        cores = os.cpu_count() or 1
        per = max(20, args.steps + args.warmup)          # one step = one new frame on every core's stream; >= 20 frames per core
        t0 = time.time()
        fps, wall = cpu_baseline_frontend(cores, per)
        ba_ref = None
        if args.workload == "all":
            from bench_ba import reference_line
            ba_ref = reference_line(args)
        line = {"impl": "reference", "metric": f"ORB extract+match fps @{CAM['w']}x{CAM['h']}", "value": round(fps, 2), "unit": "frames/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 / fps, 3), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": frontend_config(args, args.gpus),
                "note": ("the oracle port: the REAL OpenCV primitives the reference calls (cv2 4.13 FAST / resize / GaussianBlur, SIMD) + restated reference "
                         "code, one independent stream per core.  oracle/_ref holds the reference's own ORBextractor.cc / ORBmatcher.cc / Optimizer.cc + g2o as object "
                         "code, but over scalar stand-ins for OpenCV / Eigen (headers absent from the image; 175 vs 99 ms per 1241x376 extraction, 1.5 vs 9.4 LM it/s): "
                         "the faster port is the fairer baseline; the BA record carries the object-code timing too"),
                "cpu_baseline": {"value": round(fps, 2), "unit": "frames/s", "cores": cores, "kind": "port", "cpu": cpu_model(),
                                 "sample": f"{per} frames x {cores} streams (one per core), {wall:.1f} s"},
                "e2e": {"value": round(fps, 2), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        if ba_ref is not None:
            line["ba"] = ba_ref
            from bench_merge import reference_merge
            line["merge"] = reference_merge(args)
        emit(line)
        return

    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        dist.init_process_group("nccl")
    if args.workload == "ba":
        from bench_ba import bench_ba
        out = bench_ba(args, rank, world)
    elif args.workload == "merge":
        from bench_merge import bench_merge
        out = bench_merge(args, rank, world)
    else:
        out = bench_frontend(args, rank, world)
        ba = None
        if args.workload == "all":
            # second half of BASELINE.json's metric: LocalBA LM iterations/s on the 500 KF / 50k point graph (sharded over the ranks at N > 1);
            # runs before rank 0's CPU baseline so that the other ranks do not wait in a collective for it
            from bench_ba import bench_ba
            ba = bench_ba(args, rank, world)
            # configs[3]: two-map Sim3 merge + global BA of the merged map (points stay with their origin map's GPUs at N > 1)
            from bench_merge import bench_merge
            merge = bench_merge(args, rank, world)
        if rank == 0:
            fps1, wall1 = cpu_baseline_frontend(1, max(10, args.cpu_frames))
            out["cpu_baseline"] = {"value": round(fps1, 2), "unit": "frames/s", "cores": 1, "kind": "port", "cpu": cpu_model(),
                                   "sample": f"{max(10, args.cpu_frames)} frames of one stream, {wall1:.1f} s; cv2 4.13 primitives + restated reference code "
                                             "(reference front end is single-threaded per stream)"}
            if ba is not None:
                out["ba"] = ba
                out["merge"] = merge
    if rank == 0:
        emit(out)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
