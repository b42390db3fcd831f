#!/usr/bin/env python
"""bench.py — stereo frames/s of the photometric Gauss-Newton hot path at KITTI size on B200 (BASELINE.json configs[1]).

One "step" = one stereo frame for each of S independent stereo streams resident on the GPU.  Every frame runs
    FrameHessian::makeImages(left) -> TrackerAndScaler::trackNewestCoarse
and every 5th frame of a stream is a keyframe that additionally runs
    makeImages(right) -> TrackerAndScaler::optimizeScale(seed 1.0)
("frames/s = 1/(pyramid L + tracker + amortised pyramid R + scale-opt every 5th frame)"; the streams are
staggered so that every step carries S/5 keyframes).  The S streams advance in lock step: the pyramids of a step are two
kernel launches, and every Levenberg-Marquardt round of all pose trackers AND scale optimisers is ONE launch of the fused
residual / Jacobian kernel.  Workload: synthetic stereo pairs 1232x368 (KITTI 1241x376 after the calibration crop), 2000 active
points -> ~10k template pixels at level 0, 5 pyramid levels.

The steps are software-pipelined like a live system that receives image k+1 while it tracks image k: the pyramids of step k+1
are built on the session's pyramid stream while the LM rounds of step k run (K timed steps = K pyramid builds + K tracking passes).

  value  : frames/s with the raw images already resident in HBM (pyramid build + tracking + scale optimisation timed)
  e2e    : the same through the C ABI with HOST buffers: pinned-host images (capture arenas: one H2D per arena and step)
           uploaded inside the timed region, the left
           pyramid mirrored back into the reference's host layouts for the untouched DSO code that reads it (level-0 dI on
           every frame — ImmaturePoint::traceOn; all levels of dIp + absSquaredGrad on keyframes — pixel selector, BA,
           LoopHandler), poses / scales returned to the host
  roofline : fused pose residual kernel, algorithmic bytes (64 B per template point per evaluation) over the
           CUDA-event duration of every launch, measured live in a separate profiled pass of the same steps
  cpu_baseline : the CPU oracle (-O3 -march=native, the reference's SSE accumulation order) on ONE core — the reference's
           tracker / scale optimiser / pyramid are single-threaded
  --impl reference : the same CPU path on all host cores (independent frames per thread)
Multi-GPU (--gpus N under torchrun): the tracking path does not shard — N independent replicas (weak scaling, no
collective); the Scan-Context database scan is the sharded piece and is reported in the "scan_context" object of the same
JSON line (100k descriptors row-sharded over the N ranks, one NCCL all-reduce(min) of packed keys per query batch).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from direct_stereo_slam_b200 import synthetic as syn  # noqa: E402

WORKLOAD = "kitti_1232x368_tracker+scaleopt_2000pts_5lvl"
BYTES_PER_POINT = 64  # 16 B template record + 4 taps x 12 B
IDENT7 = np.array([0, 0, 0, 1, 0, 0, 0], np.float64)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """(dram bytes per launch of the pose kernel, where that figure comes from) from the committed ncu --set full capture
    (profiles/), or (None, None)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            t = json.load(open(p))
            return t.get("pose_eval_dram_bytes_per_launch"), t.get("source")
        except Exception:
            return None, None
    return None, None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, smax, reasons = [], None, set()
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------------
# workload
# ----------------------------------------------------------------------------------------------------------------------
def make_cases(n_cases, seed0=1000):
    """Distinct synthetic stereo problems (keyframe, two new left frames, right frame, active points with a scale error)."""
    cases = []
    for k in range(n_cases):
        rng = np.random.default_rng(seed0 + k)
        c = syn.make_tracking_case("kitti", seed0 + k, motion_scale=1.0, scale_error=float(rng.uniform(0.8, 1.25)))
        # a second new frame of the same keyframe (other motion) so consecutive steps of a stream see different images
        R2, t2 = syn.se3_exp_mat(c["xi_true"] * 0.6)
        c["img_new2"], _ = c["scene"].render(R2, t2, noise_seed=(seed0 + k) * 3 + 7, aff=c["aff_true"])
        c["pose_init"] = [syn.pose7(*syn.se3_exp_mat(c["xi_true"] * 0.9)), syn.pose7(*syn.se3_exp_mat(c["xi_true"] * 0.6 * 0.9))]
        cases.append(c)
    return cases


class GpuStreams:
    """S independent stereo streams of one rank (device twins of S TrackerAndScaler objects and their frames)."""

    def __init__(self, api, session, cases, n_streams, kf_every=5):
        self.api, self.s = api, session
        self.kf_every = kf_every
        cfg = cases[0]["cfg"]
        self.w, self.h = cfg["w"], cfg["h"]
        self.levels = api.pyr_levels_used(self.w, self.h)
        K = np.array([cfg["fx"], cfg["fy"], cfg["cx"], cfg["cy"]], np.float32)
        T = syn.t_stereo(cfg).reshape(-1)
        self.n = n_streams
        self.trk, self.f_new, self.f_right, self.case_of = [], [], [], []
        self.h_new, self.h_right = [], []
        ref = api.FrameHessian(session, self.w, self.h, self.levels)
        # pinned host images live in capture arenas: the left images of a step (and the right images of its keyframes) lie
        # back to back, so dslam_frame_upload_batch moves them in one host-to-device transfer each
        arena_new = session.pinned((2, n_streams, self.h, self.w))
        arena_right = session.pinned((n_streams, self.h, self.w))
        order = sorted(range(n_streams), key=lambda i: (i % kf_every, i))  # the keyframes of a step share i % kf_every
        slot_right = {i: k for k, i in enumerate(order)}
        for i in range(n_streams):
            c = cases[i % len(cases)]
            self.case_of.append(c)
            t = api.TrackerAndScaler(session, self.w, self.h, T, K, K0=K, levels=self.levels)
            ref.makeImages(c["img_ref"], host=False)
            t.setCoarseTrackingRef(ref, c["pu"], c["pv"], c["pid"], c["pw"])
            self.trk.append(t)
            fn = [api.FrameHessian(session, self.w, self.h, self.levels) for _ in range(2)]
            fr = [api.FrameHessian(session, self.w, self.h, self.levels) for _ in range(2)]  # double-buffered like the left frames
            # pinned host copies of the inputs (e2e) and pinned host mirrors of the left pyramid
            hn = [arena_new[0, i], arena_new[1, i]]
            hn[0][:] = c["img_new"]
            hn[1][:] = c["img_new2"]
            hr = arena_right[slot_right[i]]
            hr[:] = c["img_right"]
            for f in fn:
                f.alloc_host(pinned=True)
            self.f_new.append(fn)
            self.f_right.append(fr)
            self.h_new.append(hn)
            self.h_right.append(hr)
        ref.close()
        session.sync()
        self.pc_n0 = [t.ref_level(0)[0].size for t in self.trk]
        self.plans = {}

    def upload_inputs(self):
        """Raw images into HBM (outside the timed region of the device-resident measurement)."""
        for i in range(self.n):
            for v in range(2):
                self.f_new[i][v].upload(self.h_new[i][v])
            for v in range(2):
                self.f_right[i][v].upload(self.h_right[i])
        self.s.sync()

    def is_kf(self, i, k):
        return (k + i) % self.kf_every == 0

    def _plans(self, k):
        """Pointer arrays / result buffers of step k's calls, built once per (frame parity, keyframe phase)."""
        key = (k & 1, k % self.kf_every)
        if key not in self.plans:
            api, v = self.api, k & 1
            kf = [i for i in range(self.n) if self.is_kf(i, k)]
            left = [self.f_new[i][v] for i in range(self.n)]
            right = [self.f_right[i][v] for i in kf]
            self.plans[key] = {
                "kf": kf, "left": left,
                "left_fb": api.FrameBatchPlan(left, [self.h_new[i][v] for i in range(self.n)]),
                "right_fb": api.FrameBatchPlan(right, [self.h_right[i] for i in kf]) if kf else None,
                "lm": api.LmBatchPlan(self.trk, left, [self.trk[i] for i in kf], right, self.levels - 1),
                "poses": np.stack([self.case_of[i]["pose_init"][v] for i in range(self.n)]),
                "affs": np.zeros((self.n, 2)), "scales": np.ones(len(kf), np.float32)}
        return self.plans[key]

    def _prepare(self, k, e2e):
        """Inputs + pyramids of step k: (e2e: one H2D per capture arena,) two launches for all left pyramids, two for the right
        pyramids of the step's keyframes — on the session's pyramid stream, so that the LM rounds queued next (on the frames of
        the PREVIOUS step) run concurrently — and (e2e) the host mirrors start draining on the frames' copy streams."""
        P = self._plans(k)
        left = P["left"]
        mirrors = e2e is True  # e2e == "pose_only": host images in, poses / scales out, no pyramid mirrors
        if e2e:
            if mirrors:
                for f in left:
                    f.wait_host()  # the mirror this frame object produced two steps ago must have landed before it is reused
            P["left_fb"].upload()
            if P["right_fb"]:
                P["right_fb"].upload()
        P["left_fb"].build(stage_host=3 if mirrors else 0, overlap=True)
        if P["right_fb"]:
            P["right_fb"].build(overlap=True)
        if mirrors:
            for i in range(self.n):
                if self.is_kf(i, k):
                    left[i].download(wait=False)
                else:
                    left[i].download(wait=False, levels=[0], abs_grad=False)
        self.prepared = (k, e2e)

    def step(self, k, e2e):
        """One stereo frame of every stream.  Software pipeline over the steps, like a live system that receives image k+1
        while it tracks image k: the pyramids of step k+1 are built (pyramid stream) while the LM rounds of step k run, so
        every call does one pyramid build and one tracking pass — K steps = K builds + K tracking passes."""
        if getattr(self, "prepared", None) != (k, e2e):
            self._prepare(k, e2e)        # first step of a run: nothing to overlap with
        self._prepare(k + 1, e2e)        # next frames: H2D + pyramids (+ mirrors) overlap the LM rounds below
        P = self._plans(k)
        ok, poses, affs, last, rmse, scales = P["lm"].run(P["poses"], P["affs"], P["scales"])
        # e2e: the host mirrors keep draining (per-frame copy streams) while the next step computes; they are waited for when
        # their frame object is reused (two steps later) and by drain() before the clock stops
        return ok, poses, scales, rmse

    def drain(self):
        for fn in self.f_new:
            for f in fn:
                f.wait_host()

    def h2d_bytes(self):
        return int(self.n * (1 + 1.0 / self.kf_every) * self.w * self.h * 4)

    def d2h_bytes(self):
        tot = sum((self.w >> l) * (self.h >> l) for l in range(self.levels))
        per_kf = tot * 16
        per_nonkf = self.w * self.h * 12
        return int(self.n * (per_kf / self.kf_every + per_nonkf * (1 - 1.0 / self.kf_every)) + self.n * (7 * 8 + 2 * 8 + 5 * 8 + 4))


def timed_steps(streams, session, steps, warmup, e2e, dist_barrier):
    for k in range(warmup):
        streams.step(k, e2e)
    if e2e is True:
        streams.drain()
    session.sync()
    dist_barrier()
    l0 = session.launch_count()
    session.mark(0)
    t0 = time.perf_counter()
    for k in range(steps):
        streams.step(warmup + k, e2e)
    if e2e is True:
        streams.drain()
    session.mark(1)
    session.sync()
    wall_ms = (time.perf_counter() - t0) * 1e3
    dev_ms = session.elapsed_ms()
    dist_barrier()
    # the LM loop is host-sequenced: the device span (CUDA events) and the host span agree to a few microseconds; take
    # the larger so that no host-side work of a step is left outside the number
    return max(dev_ms, wall_ms), session.launch_count() - l0


# ----------------------------------------------------------------------------------------------------------------------
# CPU arms (oracle = restatement of the reference's CPU algorithm; the only place bench.py executes oracle/)
# ----------------------------------------------------------------------------------------------------------------------
class CpuStream:
    def __init__(self, orc_mod, o, case, phase=0, kf_every=5):
        cfg = case["cfg"]
        self.phase, self.kf_every = phase, kf_every
        self.o, self.case = o, case
        self.w, self.h = cfg["w"], cfg["h"]
        self.levels = orc_mod.pyr_levels_used(self.w, self.h)
        K = np.array([cfg["fx"], cfg["fy"], cfg["cx"], cfg["cy"]], np.float32)
        self.trk = o.tracker(self.w, self.h, self.levels, K, K, syn.t_stereo(cfg))
        dIp_ref, _ = o.make_images(case["img_ref"], self.levels)
        self.trk.make_coarse_depth(case["pu"], case["pv"], case["pid"], case["pw"], dIp_ref)
        self.trk.set_ref_aff(1.0, 0.0, 0.0)

    def frame(self, k):
        v = k & 1
        c = self.case
        dIp_new, _ = self.o.make_images(c["img_new"] if v == 0 else c["img_new2"], self.levels)
        self.trk.set_new_frame(dIp_new, 1.0)
        ok, pose, aff, last, flow = self.trk.track_newest_coarse(0, c["pose_init"][v], (0.0, 0.0), self.levels - 1)
        scale = None
        if (k + self.phase) % self.kf_every == 0:  # keyframe: right pyramid + scale optimisation
            dIp_r, _ = self.o.make_images(c["img_right"], self.levels)
            self.trk.set_right_frame(dIp_r)
            rmse, scale = self.trk.optimize_scale(0, 1.0, self.levels - 1)
        return ok, pose, scale


class RefCpuStream:
    """One stereo stream on the REFERENCE'S OWN CODE: TrackerAndScaler.cpp / makeImages compiled in place from /root/reference
    (oracle/ref_build.py, -O3 -march=x86-64-v3) — the prebuilt oracle/_ref/libdslam_ref_tracker_opt.so travels to the GPU box."""

    def __init__(self, orc_mod, case, phase=0, kf_every=5):
        cfg = case["cfg"]
        self.orc, self.case, self.phase, self.kf_every = orc_mod, case, phase, kf_every
        self.w, self.h = cfg["w"], cfg["h"]
        self.levels = orc_mod.pyr_levels_used(self.w, self.h)
        K = np.array([cfg["fx"], cfg["fy"], cfg["cx"], cfg["cy"]], np.float32)
        self.path = orc_mod.ReferenceTracker.OPT_PATH
        self.trk = orc_mod.ReferenceTracker(self.w, self.h, self.levels, K, K, syn.t_stereo(cfg), path=self.path)
        self.dIp_ref, _ = orc_mod.reference_make_images(case["img_ref"], self.levels, path=self.path)
        hdif = np.full(len(case["pu"]), 1e-3, np.float32)
        self.trk.set_ref(self.dIp_ref, case["pu"], case["pv"], case["pid"], hdif)

    def frame(self, k):
        v = k & 1
        c = self.case
        dIp_new, _ = self.orc.reference_make_images(c["img_new"] if v == 0 else c["img_new2"], self.levels, path=self.path)
        self.trk.set_new_frame(dIp_new, 1.0)
        ok, pose, aff, last, flow = self.trk.track_newest_coarse(c["pose_init"][v], (0.0, 0.0), self.levels - 1)
        scale = None
        if (k + self.phase) % self.kf_every == 0:
            dIp_r, _ = self.orc.reference_make_images(c["img_right"], self.levels, path=self.path)
            self.trk.set_right_frame(dIp_r)
            rmse, scale = self.trk.optimize_scale(1.0, self.levels - 1)
        return ok, pose, scale


def make_cpu_stream(orc_mod, o, case, phase, kf_every):
    """(stream, kind): the reference's own compiled source when oracle/_ref travelled here, else the oracle port."""
    if os.path.exists(orc_mod.ReferenceTracker.OPT_PATH):
        return RefCpuStream(orc_mod, case, phase, kf_every), "reference"
    return CpuStream(orc_mod, o if o is not None else orc_mod.Oracle(native=True), case, phase=phase, kf_every=kf_every), "port"


def _ref_frames(st, k0, k1):
    """Frames [k0, k1) of a CPU stream.  RefCpuStream: the whole loop on the C side (the reference's own makeImages incl. its
    new[] / delete[], trackNewestCoarse, optimizeScale; no per-frame Python or numpy work); oracle port: per-frame calls."""
    if isinstance(st, RefCpuStream):
        c = st.case
        st.trk.run_frames(k0, k1, st.phase, st.kf_every, c["img_new"], c["img_new2"], c["img_right"], np.stack(c["pose_init"]), st.levels - 1)
    else:
        for k in range(k0, k1):
            st.frame(k)


def cpu_single_core(cases, kf_every=5, budget_s=12.0):
    import oracle as orc

    st, kind = make_cpu_stream(orc, None, cases[0], 0, kf_every)
    _ref_frames(st, 0, 5)  # warm-up
    n, chunk, t0 = 0, 20, time.perf_counter()
    while True:
        _ref_frames(st, n, n + chunk)
        n += chunk
        dt = time.perf_counter() - t0
        if dt > budget_s or n >= 4000:
            break
    how = ("the reference's own TrackerAndScaler.cpp / makeImages compiled in place (oracle/_ref, -O3 -march=x86-64-v3), frame loop on the C side"
           if kind == "reference" else "oracle port (-O3 -march=native, SSE accumulation order)")
    return {"value": n / dt, "unit": "frames/s", "cores": 1, "kind": kind,
            "sample": "%d stereo frames of %s on one core, %s, %.1f s" % (n, WORKLOAD, how, dt)}


def _cpu_proc_main(idx, core, n_cases, kf_every, warmup, steps, barrier, out_q):
    """One independent stereo stream in its own PROCESS pinned to one core — the reference is single-threaded per rig, and
    threads of one process contend on the allocator / page-fault path of the ~10 MB per-frame pyramids (round-1 harness)."""
    try:
        if core is not None:
            os.sched_setaffinity(0, {core})
    except OSError:
        pass
    import oracle as orc

    cases = make_cases(1, seed0=1000 + idx % n_cases)
    st, kind = make_cpu_stream(orc, None, cases[0], idx, kf_every)
    barrier.wait()
    _ref_frames(st, 0, warmup)
    barrier.wait()
    t0 = time.perf_counter()
    _ref_frames(st, warmup, warmup + steps)
    t1 = time.perf_counter()
    out_q.put((idx, t0, t1, kind))


def cpu_procs(n_procs, steps, warmup, n_cases=4, kf_every=5):
    """(frames/s, span seconds, kind) of n_procs independent stereo streams, one process per core; the span is
    max(end) - min(start) over the processes (CLOCK_MONOTONIC is system wide)."""
    import multiprocessing as mp

    ctx = mp.get_context("fork")
    cores = sorted(os.sched_getaffinity(0))
    barrier, q = ctx.Barrier(n_procs), ctx.Queue()
    procs = [ctx.Process(target=_cpu_proc_main, args=(i, cores[i % len(cores)], n_cases, kf_every, warmup, steps, barrier, q)) for i in range(n_procs)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=1800) for _ in procs]
    for pr in procs:
        pr.join()
    span = max(r[2] for r in res) - min(r[1] for r in res)
    return n_procs * steps / span, span, res[0][3]


def cpu_all_cores(steps, warmup, n_cases=4, kf_every=5):
    """Reference arm: every usable host core runs its own stream.  Also reports the 1-process rate measured the same way,
    so that harness throttling would be visible as an efficiency well below 1."""
    n = len(os.sched_getaffinity(0))
    fps1, _, _ = cpu_procs(1, max(steps, 20), warmup, n_cases, kf_every)
    fps, span, kind = cpu_procs(n, steps, warmup, n_cases, kf_every)
    return fps, span, kind, n, {"procs_1_fps": fps1, "procs_%d_fps" % n: fps, "efficiency_vs_linear": fps / (n * fps1)}


# ----------------------------------------------------------------------------------------------------------------------
# Scan-Context shard (the only piece that shards): 100k descriptors over the ranks, query batch of 32
# ----------------------------------------------------------------------------------------------------------------------
def bench_scan_context(api, session, rank, world, dist, n_db=100_000, batches=(1, 32, 256), reps=20):
    """100k descriptors row-sharded over the ranks; per query batch: scan + exact re-score (+ one NCCL all-reduce(min))."""
    sig, key = syn.make_sc_database(n_db, 2024)
    rows = api.shard_rows(n_db, world, rank)
    db = api.ScanContextDB(session, len(rows) + 8)
    db.add(key[rows], sig[rows], global_ids=rows)
    if world > 1:
        import torch

        ident = [api.ScanContextDB.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ident, src=0)
        db.attach_comm(ident[0], world, rank)
    peak, _ = measured_peak()
    mode = db.exchange_mode()
    out = {"db_rows": n_db, "rows_per_gpu": int(len(rows)),
           "exchange": {"nvlink-mailbox": "system-scope atomic-min of packed (distance, id) keys into every rank's HBM mailbox over NVLink, fused into the re-score kernel",
                        "nccl": "ncclAllReduce(min, uint64 x Q)", "none": "none (one GPU)"}[mode], "batches": {}}
    full = None
    for nq in batches:
        qs, qk, truth = syn.make_sc_queries(sig, key, nq, 77)
        for _ in range(3):
            idx, diff = db.query(qs)
        lat, scan = [], []
        for _ in range(reps):
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            idx, diff = db.query(qs)
            lat.append((time.perf_counter() - t0) * 1e3)
            scan.append(db.last_scan_ms())
        lat_ms, scan_ms = float(np.median(lat)), float(np.median(scan))
        if world > 1:
            import torch

            t = torch.tensor([lat_ms, scan_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            lat_ms, scan_ms = float(t[0]), float(t[1])
        known = truth >= 0
        gbs = len(rows) * (db.n_cells + db.n_rings) * 4 / (scan_ms * 1e-3) / 1e9
        out["batches"]["Q%d" % nq] = {"query_latency_ms": lat_ms, "scan_kernel_ms": scan_ms, "scan_gbs_per_gpu": gbs, "scan_frac_of_hbm_peak": gbs / peak,
                                      "revisits_found": int(np.sum(idx[known] == truth[known])), "revisits": int(known.sum())}
        if world > 1 and rank == 0:
            # parity of the sharded answer: the same batch against an UNSHARDED copy of the database on this GPU (non-collective
            # call) must give the same argmin and the same distance bits
            if full is None:
                full = api.ScanContextDB(session, n_db + 8)
                full.add(key, sig)
            dref, iref = api.unpack_key(full.query_keys(qs))
            out["batches"]["Q%d" % nq]["equals_unsharded"] = bool(np.array_equal(iref, idx) and np.array_equal(dref.view(np.uint32), diff.view(np.uint32)))
    if full is not None:
        full.close()
    db.close()
    return out


# ----------------------------------------------------------------------------------------------------------------------
# single-stream latency (what ONE stereo rig at 30 Hz sees; src/main.cpp:212-265 -> src/FrontEnd.cpp:585-686)
# ----------------------------------------------------------------------------------------------------------------------
def bench_latency(api, session, case, kf_every=5, frames=60):
    """One stereo stream, frame after frame, in FrontEnd::addActiveStereoFrame's order: makeImages(left) [pinned host image ->
    device pyramid -> level-0 dI mirrored back to the host for ImmaturePoint::traceOn; all levels + absSquaredGrad on keyframes],
    trackNewCoarse with the 83-hypothesis list of src/FrontEnd.cpp:147-180 (dslam_track_new_coarse: the first hypothesis is
    evaluated alone and normally accepted, :244-246), and on every kf_every-th frame makeImages(right) + optimizeScale(1.0).
    Host clock around each call; median over the frames."""
    cfg = case["cfg"]
    w, h = cfg["w"], cfg["h"]
    levels = api.pyr_levels_used(w, h)
    K = np.array([cfg["fx"], cfg["fy"], cfg["cx"], cfg["cy"]], np.float32)
    trk = api.TrackerAndScaler(session, w, h, syn.t_stereo(cfg).reshape(-1), K, K0=K, levels=levels)
    ref = api.FrameHessian(session, w, h, levels)
    ref.makeImages(case["img_ref"], host=False)
    trk.setCoarseTrackingRef(ref, case["pu"], case["pv"], case["pid"], case["pw"])
    fl = [api.FrameHessian(session, w, h, levels) for _ in range(2)]
    fr = api.FrameHessian(session, w, h, levels)
    for f in fl:
        f.alloc_host(pinned=True)
    imgs = [session.pinned((h, w)), session.pinned((h, w))]
    imgs[0][:] = case["img_new"]
    imgs[1][:] = case["img_new2"]
    img_r = session.pinned((h, w))
    img_r[:] = case["img_right"]
    ident = IDENT7
    tries = []
    for v in range(2):
        xi = case["xi_true"] * (1.0 if v == 0 else 0.6) * 0.9
        tries.append(syn.frontend_pose_tries(case["pose_init"][v], syn.pose7(*syn.se3_exp_mat(xi * 2)), syn.pose7(*syn.se3_exp_mat(xi * 0.5)), ident))
    session.sync()
    T = {"make_images_left_call": [], "track_new_coarse": [], "mirror_wait": [], "frame_nonkf": [], "make_images_right+optimize_scale": [], "frame_kf": []}
    last = np.full(5, 100.0)
    tries_used = []
    for k in range(-5, frames):
        v, kf = k & 1, (k % kf_every == 0)
        f = fl[v]
        t0 = time.perf_counter()
        f.upload(imgs[v])
        f.build()
        if kf:
            f.download(wait=False)
        else:
            f.download(wait=False, levels=[0], abs_grad=False)
        t1 = time.perf_counter()
        r = trk.trackNewCoarse(f, tries[v], (0.0, 0.0), levels - 1, last)
        t2 = time.perf_counter()
        f.wait_host()
        t3 = time.perf_counter()
        if kf:
            fr.upload(img_r)
            fr.build()
            trk.optimizeScale(fr, 1.0, levels - 1)
        t4 = time.perf_counter()
        last = np.where(np.isfinite(r["achievedRes"]), r["achievedRes"], last)
        if k < 0:
            continue
        tries_used.append(r["tryIterations"])
        T["make_images_left_call"].append(t1 - t0)
        T["track_new_coarse"].append(t2 - t1)
        T["mirror_wait"].append(t3 - t2)
        if kf:
            T["make_images_right+optimize_scale"].append(t4 - t3)
            T["frame_kf"].append(t4 - t0)
        else:
            T["frame_nonkf"].append(t3 - t0)
    # the pyramid alone (upload + build, device complete), and the full makeImages contract (all levels mirrored, blocking)
    pyr, full = [], []
    for k in range(20):
        t0 = time.perf_counter()
        fl[0].upload(imgs[0])
        fl[0].build()
        session.sync()
        pyr.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        fl[1].upload(imgs[1])
        fl[1].build()
        fl[1].download(wait=True)
        full.append(time.perf_counter() - t0)
    med = {k: float(np.median(v)) * 1e3 for k, v in T.items() if v}
    out = {"streams": 1, "unit": "ms", "frames": frames, "keyframe_every": kf_every, "hypotheses_offered": int(len(tries[0])),
           "hypotheses_tried_mean": float(np.mean(tries_used)), "median_ms": med,
           "pyramid_h2d+build_ms": float(np.median(pyr)) * 1e3, "make_images_all_levels_mirrored_ms": float(np.median(full)) * 1e3,
           "frame_ms_amortised": (float(np.sum(T["frame_nonkf"])) + float(np.sum(T["frame_kf"]))) / frames * 1e3}
    out["frames_per_sec_single_stream"] = 1e3 / out["frame_ms_amortised"]
    for x in [trk, ref, fr] + fl:
        x.close()
    return out


def cpu_latency(cases, kf_every=5, frames=60):
    """The same per-frame sequence on the reference's own code (one core), split into its phases (reft_run_frames)."""
    import oracle as orc

    st, kind = make_cpu_stream(orc, None, cases[0], 0, kf_every)
    if not isinstance(st, RefCpuStream):
        return None
    _ref_frames(st, 0, 5)
    _ref_frames(st, 0, frames)
    ph = st.trk.last_phase_s
    nkf = len([k for k in range(frames) if k % kf_every == 0])
    return {"kind": kind, "cores": 1, "unit": "ms", "make_images_left": ph[0] / frames * 1e3, "track_newest_coarse": ph[1] / frames * 1e3,
            "make_images_right+optimize_scale": ph[2] / max(nkf, 1) * 1e3, "frame_ms_amortised": float(ph.sum()) / frames * 1e3}


# ----------------------------------------------------------------------------------------------------------------------
# BASELINE configs[3]: synthetic 1920x1200, 8000 active points, B pose hypotheses per call (the HBM-roofline run)
# ----------------------------------------------------------------------------------------------------------------------
def bench_sweep_config3(api, session, peak, batches=(1, 8, 64, 512), reps=10):
    c = syn.make_tracking_case("synth1920", 42, with_right=False)
    cfg = c["cfg"]
    w, h = cfg["w"], cfg["h"]
    levels = api.pyr_levels_used(w, h)
    K = np.array([cfg["fx"], cfg["fy"], cfg["cx"], cfg["cy"]], np.float32)
    trk = api.TrackerAndScaler(session, w, h, syn.t_stereo(cfg).reshape(-1), K, K0=K, levels=levels)
    ref = api.FrameHessian(session, w, h, levels)
    ref.makeImages(c["img_ref"], host=False)
    pcn = trk.setCoarseTrackingRef(ref, c["pu"], c["pv"], c["pid"], c["pw"])
    f = api.FrameHessian(session, w, h, levels)
    f.makeImages(c["img_new"], host=False)
    rng = np.random.default_rng(42)
    out = {"workload": "synthetic_1920x1200_8000pts", "template_points_lvl0": int(pcn[0]), "levels": levels, "batches": {}}
    session.profile(True)
    for B in batches:
        hyp = np.stack([syn.pose7(*syn.se3_exp_mat(c["xi_true"] * (1 + rng.uniform(-0.02, 0.02, 6)))) for _ in range(B)])
        aff = np.tile(np.array(c["aff_true"]), (B, 1)) * (1 + rng.uniform(-0.02, 0.02, (B, 2)))
        for lvl in (0,):
            for _ in range(3):
                trk.calcResAndGSPose(f, lvl, hyp, aff)
            session.profile_read()
            t0 = time.perf_counter()
            for _ in range(reps):
                trk.calcResAndGSPose(f, lvl, hyp, aff)
            call_us = (time.perf_counter() - t0) / reps * 1e6
            r = session.profile_read()["pose"]
            gbs = BYTES_PER_POINT * r["points"] / (r["ms"] * 1e-3) / 1e9
            out["batches"]["B%d" % B] = {"hypotheses": B, "launches_per_call": r["launches"] // reps, "kernel_us_per_call": r["ms"] * 1e3 / reps,
                                         "call_us_host_clock": call_us, "points_per_call": r["points"] // reps, "achieved_gbs": gbs, "frac": gbs / peak}
    session.profile(False)
    for x in (trk, ref, f):
        x.close()
    return out


def parity_check(api, session, case):
    """Checker leg (the one other place bench.py runs oracle/): the GPU's trackNewestCoarse on scene 0 against the oracle in
    its fp64-accumulating mode (the parity target) and in its reference-faithful mode 0 (4-lane x 3-tier fp32
    SSE accumulation, bit-identical to the reference's own compiled TrackerAndScaler.cpp).  noise_floor = the distance of the
    per-iteration increments to the mode-0 trace where the accept / reject sequences coincide: the reference's own
    summation noise, which no fp64-accumulating implementation can be closer to."""
    import oracle as orc

    def rel(a, b):
        return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))

    o = orc.Oracle()
    cfg = case["cfg"]
    w, h = cfg["w"], cfg["h"]
    levels = orc.pyr_levels_used(w, h)
    K = np.array([cfg["fx"], cfg["fy"], cfg["cx"], cfg["cy"]], np.float32)
    T = syn.t_stereo(cfg)
    d_ref, _ = o.make_images(case["img_ref"], levels)
    d_new, _ = o.make_images(case["img_new"], levels)
    ot = o.tracker(w, h, levels, K, K, T)
    ot.make_coarse_depth(case["pu"], case["pv"], case["pid"], case["pw"], d_ref)
    ot.set_ref_aff(1.0, 0.0, 0.0)
    ot.set_new_frame(d_new, 1.0)
    init = case["pose_init"][0]
    r1 = ot.track_newest_coarse(1, init, (0.0, 0.0), levels - 1)
    t1 = ot.trace()
    r0 = ot.track_newest_coarse(0, init, (0.0, 0.0), levels - 1)
    t0 = ot.trace()
    trk = api.TrackerAndScaler(session, w, h, T.reshape(-1), K, K0=K, levels=levels)
    ref = api.FrameHessian(session, w, h, levels)
    ref.makeImages(case["img_ref"], host=False)
    trk.setCoarseTrackingRef(ref, case["pu"], case["pv"], case["pid"], case["pw"])
    f = api.FrameHessian(session, w, h, levels)
    f.makeImages(case["img_new"], host=False)
    ok, pose, aff, last = trk.trackNewestCoarse(f, init, (0.0, 0.0), levels - 1)
    tg = trk.trace()
    for x in (trk, ref, f):
        x.close()

    def inc_dist(ta, tb):
        n, worst, errs = 0, 0.0, []
        for a, b in zip(ta, tb):
            if not np.array_equal(a[:3], b[:3]):
                break
            n += 1
            if b[1] >= 0:
                errs.append(rel(a[7:15], b[7:15]))
        return n, (max(errs) if errs else 0.0), (float(np.median(errs)) if errs else 0.0)

    n1, w1, m1 = inc_dist(tg, t1)
    n0, w0, m0 = inc_dist(tg, t0)
    return {"workload": WORKLOAD + " scene 0, trackNewestCoarse from pose_init", "lm_rows": int(len(tg)),
            "vs_fp64_oracle": {"rows_coinciding": n1, "rows": int(len(t1)), "inc_rel_worst": w1, "final_pose_rel": rel(pose, r1[1]), "ok_equal": bool(ok == r1[0]),
                               "tolerance": 1e-5},
            "noise_floor": {"what": "vs the oracle's reference-faithful mode 0 (fp32 SSE accumulation order, bit-identical to the compiled reference)",
                            "rows_coinciding": n0, "rows": int(len(t0)), "inc_rel_worst": w0, "inc_rel_median": m0, "final_pose_rel": rel(pose, r0[1]),
                            "ok_equal": bool(ok == r0[0])}}


def bind_to_gpu_numa_node(local_rank):
    """Multi-GPU runs: keep this rank's host threads (LM groups, copy threads) and the pinned arenas they first touch on the CPUs
    that are local to its GPU (sysfs local_cpulist of the GPU's PCI function), when that is a proper subset of the CPUs the
    process may use.  Returns what was done, for the JSON line."""
    info = {"numa_node": None, "local_cpus": None, "bound": False}
    try:
        out = subprocess.run(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True,
                             timeout=20).stdout.strip()
        bdf = out.lower()
        if bdf.startswith("0000"):
            bdf = bdf[4:]  # nvidia-smi prints an 8-digit domain, sysfs a 4-digit one
        base = "/sys/bus/pci/devices/" + bdf
        info["numa_node"] = int(open(base + "/numa_node").read())
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        local = cpus & allowed
        info["local_cpus"] = len(local)
        if local and local != allowed:
            os.sched_setaffinity(0, local)
            info["bound"] = True
    except Exception as e:  # sysfs layout differs / container hides it: run unbound, say so
        info["error"] = str(e)[:100]
    return info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=512, help="independent stereo streams per GPU advanced in lock step")
    ap.add_argument("--keyframe-every", type=int, default=5, help="every k-th frame of a stream also builds the right pyramid and optimises the scale")
    ap.add_argument("--cases", type=int, default=4, help="distinct synthetic scenes (streams cycle through them)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-scan-context", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the single-stream latency, the config-3 sweep and the parity check")
    args = ap.parse_args()
    warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        fps, dt, kind, threads, per_core = cpu_all_cores(args.steps, warmup, min(args.cases, 4), args.keyframe_every)
        line = {"metric": "stereo_frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": warmup,
                "ms_per_step": dt * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "impl": "reference",
                "config": {"workload": WORKLOAD, "frames_per_step": threads, "keyframe_every": args.keyframe_every,
                           "note": "the reference's own TrackerAndScaler.cpp / FrameHessian::makeImages source compiled in place against Eigen/Sophus "
                                   "stand-ins (oracle/ref_build.py) when kind == reference, else the oracle port; one independent stereo stream per "
                                   "host core, each in its own process pinned to that core, frame loop on the C side"},
                "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind,
                                 "sample": "%d processes x %d stereo frames of %s" % (threads, args.steps, WORKLOAD), "per_core_scaling": per_core},
                "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    from direct_stereo_slam_b200 import api

    dist = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL's version / debug banner must not share stdout with the JSON line
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()

    binding = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    session = api.Session(local_rank)
    cases = make_cases(args.cases, seed0=1000 + 16 * rank)
    streams = GpuStreams(api, session, cases, args.streams, kf_every=args.keyframe_every)

    # ---- value: device-resident inputs ------------------------------------------------------------------------------------
    streams.upload_inputs()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # samples every 100 ms across both timed regions (device-resident and e2e)
    session.host_times()
    ms, launches = timed_steps(streams, session, args.steps, warmup, False, barrier)
    host_times = session.host_times()
    # ---- e2e: host buffers through the C ABI ---------------------------------------------------------------------------------
    ms_e2e, _ = timed_steps(streams, session, args.steps, warmup, True, barrier)
    # ---- the path's own end-to-end cost: host images in, poses / scales out, no host mirrors of the pyramid ------------------
    ms_e2e_pose, _ = timed_steps(streams, session, args.steps, warmup, "pose_only", barrier)
    clocks = sampler.stop() if rank == 0 else None
    # ---- roofline: per-launch CUDA-event timing of the fused pose kernel over the same steps -----------------------------------
    session.profile(True)
    for k in range(min(args.steps, 5)):
        streams.step(k, False)
    prof = session.profile_read()
    counters = [t.counters() for t in streams.trk]
    # the same kernel with the GPU to itself: one launch of 128 pose hypotheses of one stream's level-0 template (1.27 M points),
    # CUDA events around the launch — what the kernel does when it is not queueing behind 7 other lanes and the pyramid stream
    session.sync()
    rng = np.random.default_rng(5)
    hyp = np.tile(streams.case_of[0]["pose_init"][0], (128, 1))
    hyp[:, 4:] += rng.normal(0, 0.01, (128, 3))
    hyp_aff = rng.normal(0, [0.01, 1.0], (128, 2))
    for _ in range(3):
        streams.trk[0].calcResAndGSPose(streams.f_new[0][0], 0, hyp, hyp_aff)
    session.profile_read()
    for _ in range(10):
        streams.trk[0].calcResAndGSPose(streams.f_new[0][0], 0, hyp, hyp_aff)
    alone = session.profile_read()["pose"]
    # the scale flavour of the kernel alone: 128 scales of one stream's level-0 template against its right pyramid
    sc_vals = np.linspace(0.8, 1.25, 128).astype(np.float32)
    for _ in range(3):
        streams.trk[0].calcResAndGSScale(streams.f_right[0][0], 0, sc_vals)
    session.profile_read()
    for _ in range(10):
        streams.trk[0].calcResAndGSScale(streams.f_right[0][0], 0, sc_vals)
    alone_scale = session.profile_read()["scale"]
    session.profile(False)
    # the pyramid kernels alone: all left pyramids of a step (two launches per 64 frames) on the session stream, CUDA events
    P0 = streams._plans(0)
    for _ in range(2):
        P0["left_fb"].build(stage_host=0, overlap=False)
    session.sync()
    session.mark(0)
    for _ in range(3):
        P0["left_fb"].build(stage_host=0, overlap=False)
    session.mark(1)
    session.sync()
    pyr_ms = session.elapsed_ms() / 3
    pyr_bytes = args.streams * (4 * streams.w * streams.h + 16 * sum((streams.w >> l) * (streams.h >> l) for l in range(streams.levels)))
    latency = sweep3 = None
    if rank == 0 and not args.no_extras:
        latency = bench_latency(api, session, cases[0], args.keyframe_every)
        sweep3 = bench_sweep_config3(api, session, measured_peak()[0])

    if dist is not None:
        import torch

        t = torch.tensor([ms, ms_e2e, ms_e2e_pose], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e, ms_e2e_pose = float(t[0]), float(t[1]), float(t[2])
        tl = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(tl)
        launches = int(tl[0])

    sc = None
    if not args.no_scan_context:
        sc = bench_scan_context(api, session, rank, world, dist)

    if rank == 0:
        frames = args.streams * world * args.steps
        peak, peak_src = measured_peak()
        p = max((prof["pose"], prof["mixed"]), key=lambda d: d["ms"])
        p = dict(launches=prof["pose"]["launches"] + prof["mixed"]["launches"], ms=prof["pose"]["ms"] + prof["mixed"]["ms"],
                 points=prof["pose"]["points"] + prof["mixed"]["points"])
        achieved = BYTES_PER_POINT * p["points"] / (p["ms"] * 1e-3) / 1e9 if p["ms"] > 0 else 0.0
        line = {"metric": "stereo_frames_per_sec", "value": frames / (ms * 1e-3), "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "streams_per_gpu": args.streams, "frames_per_step": args.streams * world,
                           "template_points_lvl0": int(np.mean(streams.pc_n0)), "keyframe_every": args.keyframe_every,
                           "l2": "working set ~%d MB per step > 126 MB L2 (every stream has its own pyramids)" % int(args.streams * 12 * (1 + 1.0 / args.keyframe_every)),
                           "multi_gpu": "replicas only (tracking does not shard); scan_context is the sharded piece",
                           "timing": "max(CUDA events on the session stream, host clock) over the K steps, max over ranks"},
                "e2e": {"value": frames / (ms_e2e * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": streams.h2d_bytes() * world,
                        "d2h_bytes_per_step": streams.d2h_bytes() * world,
                        "note": "contract complete: every frame's level-0 dI (and on keyframes all levels + absSquaredGrad) is mirrored into the "
                                "reference's host layouts for the untouched DSO code that reads it; PCIe-bound (see e2e_pose_only and DESIGN.md 10)"},
                "e2e_pose_only": {"value": frames / (ms_e2e_pose * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": streams.h2d_bytes() * world,
                                  "d2h_bytes_per_step": args.streams * world * (7 * 8 + 2 * 8 + 5 * 8 + 4),
                                  "note": "host images in, poses / affine / residuals / scales out; no host mirror of the pyramid"},
                "gpu_launches": launches,
                "roofline": {"bound": "hbm", "kernel": "eval_kernel (fused calcRes*+calcGSSSE* of all pose / scale items of an LM round)", "achieved": achieved, "peak": peak,
                             "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic()[0], "traffic_source": ncu_traffic()[1], "peak_source": peak_src,
                             "launches_timed": p["launches"], "avg_launch_us": (p["ms"] * 1e3 / p["launches"]) if p["launches"] else None,
                             "points_per_launch": (p["points"] / p["launches"]) if p["launches"] else None,
                             "scale_kernel_gbs": (BYTES_PER_POINT * prof["scale"]["points"] / (prof["scale"]["ms"] * 1e-3) / 1e9)
                             if prof["scale"]["ms"] > 0 else None,
                             "note": "achieved / frac are the launches of the timed configuration: 8 LM lanes and the pyramid stream share the GPU, "
                                     "so the event-timed duration of a launch is mostly queueing; kernel_alone is the same kernel in one launch of "
                                     "128 hypotheses x the level-0 template with the GPU to itself",
                             "kernel_alone": ({"launches_timed": alone["launches"], "points_per_launch": alone["points"] / alone["launches"],
                                               "avg_launch_us": alone["ms"] * 1e3 / alone["launches"],
                                               "achieved": BYTES_PER_POINT * alone["points"] / (alone["ms"] * 1e-3) / 1e9,
                                               "frac": BYTES_PER_POINT * alone["points"] / (alone["ms"] * 1e-3) / 1e9 / peak}
                                              if alone["launches"] and alone["ms"] > 0 else None),
                             "scale_kernel_alone": ({"launches_timed": alone_scale["launches"], "points_per_launch": alone_scale["points"] / alone_scale["launches"],
                                                     "avg_launch_us": alone_scale["ms"] * 1e3 / alone_scale["launches"],
                                                     "achieved": BYTES_PER_POINT * alone_scale["points"] / (alone_scale["ms"] * 1e-3) / 1e9,
                                                     "frac": BYTES_PER_POINT * alone_scale["points"] / (alone_scale["ms"] * 1e-3) / 1e9 / peak}
                                                    if alone_scale["launches"] and alone_scale["ms"] > 0 else None),
                             "pyramid": {"kernels": "downsample_chain_kernel + gradient_kernel, all %d left pyramids of a step" % args.streams,
                                         "bytes_per_frame": pyr_bytes // args.streams, "ms_per_step": pyr_ms,
                                         "achieved": pyr_bytes / (pyr_ms * 1e-3) / 1e9, "frac": pyr_bytes / (pyr_ms * 1e-3) / 1e9 / peak,
                                         "note": "algorithmic bytes = read 4*P0 + write 16*sum(P_l) per frame (SURVEY.md 8d)"},
                             "aggregate": {"eval_bytes_per_step": BYTES_PER_POINT * p["points"] / min(args.steps, 5),
                                           "pyramid_bytes_per_step": pyr_bytes * (1 + 1.0 / args.keyframe_every),
                                           "gbs": (BYTES_PER_POINT * p["points"] / min(args.steps, 5) + pyr_bytes * (1 + 1.0 / args.keyframe_every)) / (ms / args.steps * 1e-3) / 1e9,
                                           "note": "algorithmic bytes of all kernels of a step / ms_per_step of the timed run (the lanes overlap, so per-launch "
                                                   "event times do not add up to the step)"},
                             "sweep": sweep3},
                "clocks": clocks,
                "host": {"cpus": len(os.sched_getaffinity(0)), "ranks": world, "numa_binding": binding},
                "host_ms_per_step": {k: (v / (args.steps + warmup) if k != "launches" else v) for k, v in host_times.items()},
                "lm": {"evals_per_frame": float(np.sum([c["evals"] for c in counters])) / (args.streams * (2 * warmup + 2 * args.steps + min(args.steps, 5))),
                       "note": "fused residual+Jacobian evaluations (pose + scale LM rounds) per stereo frame"}}
        line["roofline"]["frac_aggregate"] = line["roofline"]["aggregate"]["gbs"] / peak
        if sc is not None:
            line["scan_context"] = sc
        if latency is not None:
            line["latency"] = {"gpu": latency}
        if not args.no_cpu_baseline and world >= 1:
            line["cpu_baseline"] = cpu_single_core(cases, args.keyframe_every)
            if latency is not None:
                line["latency"]["cpu_reference"] = cpu_latency(cases, args.keyframe_every)
            if not args.no_extras:
                line["parity"] = parity_check(api, session, cases[0])
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
