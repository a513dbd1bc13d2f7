#!/usr/bin/env python
"""
bench.py - headline benchmark of the Pan-tilt-zoom-SLAM hot path on B200.

Metric (BASELINE.json): residual + Jacobian + normal-equation assembly, OBSERVATIONS PER SECOND, on config 3
(keyframe BA, 256 keyframes x 100k ray landmarks x 2M observations); BA LM iterations/s is reported beside it
(`lm_iters_per_s`).  A step = one fused pass (ptzba_ba_normal_equations) over the whole observation list.

  value     device-resident: x and all observation arrays already in HBM, K passes timed with CUDA events on the
            launching stream; the passes rotate over R replicas of the problem so that every pass streams its
            observations from HBM, not from the 126 MB L2 (config: l2 = "rotating replicas").
  e2e       the same pass through the C-ABI with HOST (pinned) buffers: H2D of x, the kernels, D2H of the residual
            vector, the U/g_c/V/g_l blocks and the cost, all inside the timed region.
  roofline  algorithmic bytes of the fused kernel (40 B/obs + 56 B/landmark + 96 B/keyframe) divided
            by that kernel's own average duration (CUDA events bracketing each launch) over the measured HBM peak.
  cpu_baseline  oracle/ptz_oracle_c.c (plain-C port of the reference algorithm, all host threads) on the same workload.

`--impl reference` times only that CPU port (the reference itself is pure Python and cannot travel to the GPU box).
Multi-GPU (torchrun, one rank per GPU): ONE problem of the named workload, its observations sharded by keyframe
(dist.shard_by_keyframe); every rank runs the fused pass on its shard and the blocks of the landmarks observed by more than
one rank are summed with ncclAllReduce INSIDE the timed region (strong scaling: `value` = observations of the whole
problem / max-over-ranks time).  After the timed loop every rank checks its blocks against a single-GPU pass over the whole
problem (`parity`), and the LM iteration / solve of the same problem is timed with the landmark-partitioned distributed solver.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    "cfg3": dict(n_kf=256, n_lm=100000, n_obs=2000000, seed=1003, pan_sweep=None),
    "cfg5": dict(n_kf=1024, n_lm=1000000, n_obs=20000000, seed=1005, pan_sweep=40.0),
    "small": dict(n_kf=32, n_lm=5000, n_obs=60000, seed=1001, pan_sweep=None),
}
BYTES_PER_OBS, BYTES_PER_LM, BYTES_PER_KF = 40, 56, 96


def algorithmic_bytes(n_obs, n_lm, n_kf):
    return n_obs * BYTES_PER_OBS + n_lm * BYTES_PER_LM + n_kf * BYTES_PER_KF


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed regions run (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_workload(name, rank=0):
    import ptz_slam_b200  # noqa: F401
    from ptz_slam_b200 import synth
    w = WORKLOADS[name]
    return synth.make_flat_ba(w["n_kf"], w["n_lm"], w["n_obs"], seed=w["seed"] + 17 * rank, pan_sweep=w["pan_sweep"],
                              id_order=os.environ.get("PTZBA_BENCH_ID_ORDER", "first_keyframe"))


# -------------------------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline
# -------------------------------------------------------------------------------------------------------------------
def time_cpu_port(fb, steps, warmup, budget_s=20.0):
    """Plain-C port (oracle/ptz_oracle_c.c), all host threads, full fused pass per step on the same workload.
    Steps are bounded so the run stays within `budget_s` seconds of CPU work."""
    from oracle import c_port, ptz_oracle as O
    from ptz_slam_b200 import synth
    poses, rays = O.ba_unpack(fb.x0(), fb.n_pose, fb.ptz_init[0])
    threads = c_port.max_threads()
    t0 = time.perf_counter()
    c_port.ba_fused(poses, rays, fb.cam_idx, fb.lm_idx, fb.obs_xy, synth.PP_U, synth.PP_V, n_threads=threads)
    one = time.perf_counter() - t0
    k = max(1, min(steps, int(budget_s / max(one, 1e-6))))
    n_warm = max(0, min(warmup - 1, int(0.25 * budget_s / max(one, 1e-6))))      # the probe pass above was the first warm-up pass
    for _ in range(n_warm):
        c_port.ba_fused(poses, rays, fb.cam_idx, fb.lm_idx, fb.obs_xy, synth.PP_U, synth.PP_V, n_threads=threads)
    t0 = time.perf_counter()
    for _ in range(k):
        c_port.ba_fused(poses, rays, fb.cam_idx, fb.lm_idx, fb.obs_xy, synth.PP_U, synth.PP_V, n_threads=threads)
    dt = time.perf_counter() - t0
    return dict(value=fb.n_obs * k / dt, steps=k, warmup=n_warm + 1, ms_per_step=1e3 * dt / k, cores=threads)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fb = make_workload(args.workload)
    r = time_cpu_port(fb, args.steps, args.warmup, budget_s=60.0)
    sample = "%d full fused passes over all %d observations of %s" % (r["steps"], fb.n_obs, args.workload)
    line = {
        "impl": "reference", "metric": "ba_residual_jacobian_normal_eq_obs_per_s", "value": r["value"], "unit": "obs/s",
        "n_gpus": args.gpus, "steps": r["steps"], "warmup": r["warmup"], "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s keyframe BA: %d keyframes x %d ray landmarks x %d observations%s" %
                   (args.workload, fb.n_pose, fb.n_landmark, fb.n_obs,
                    " per GPU" if args.gpus == 1 else ", ONE problem sharded by keyframe over %d GPUs" % args.gpus),
                   "parallelism": "host threads (plain-C port of the reference's pass, oracle/ptz_oracle_c.c)",
                   "step": "one fused residual+Jacobian+normal-equation pass over the whole workload"},
        "cpu_baseline": {"value": r["value"], "unit": "obs/s", "cores": r["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": r["value"], "unit": "obs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def fp64_peak():
    """Measured FP64 peaks of this pool's B200 (scripts/fp64_peak.cu, profiles/r2_fp64_peak.json); MEASURED_PEAKS.json has no FP64 figure."""
    p = os.path.join(ROOT, "profiles", "r2_fp64_peak.json")
    try:
        j = json.load(open(p))
        return {"dfma_tflops": j["dfma_tflops"], "dmma_tflops": j["dmma_m8n8k4_tflops"], "source": "profiles/r2_fp64_peak.json (measured)"}
    except Exception:
        return {"dfma_tflops": 34.0, "dmma_tflops": 37.0, "source": "fallback (round-2 measurement)"}


def verbatim_reference_record():
    """The UNMODIFIED reference timed on host cores (scripts/time_verbatim_reference.py -> profiles/r2_verbatim_reference_cpu.json).
    /root/reference does not travel to the GPU box, so these figures come from the authoring container and are quoted, not
    re-measured: a reported baseline beside `cpu_baseline` (which IS measured on this box), never an input of any ratio."""
    p = os.path.join(ROOT, "profiles", "r2_verbatim_reference_cpu.json")
    try:
        j = json.load(open(p))
        res = max(j["compute_residual"], key=lambda r: r["observations"])
        ls = max(j["least_squares"], key=lambda r: r["observations"])
        ekf = {str(r["rays"]): r["verbatim_reference"]["frames_per_s"] for r in j["ekf_update"]}
        return {"where": "authoring container, %d host CPUs, NOT this box (%s)" % (j["host"]["cpus"], "profiles/r2_verbatim_reference_cpu.json"),
                "compute_residual_obs_per_s": res["verbatim_reference"]["obs_per_s"],
                "compute_residual_sample": "%d keyframes x %d landmarks x %d observations, one core (pure-Python loops)" % (res["keyframes"], res["landmarks"], res["observations"]),
                "least_squares_s_per_lm_iteration": ls["verbatim_reference"]["s_per_lm_iteration"],
                "least_squares_sample": "%d keyframes x %d landmarks x %d observations, %d parameters: %.1f s, nfev %d" %
                                        (ls["keyframes"], ls["landmarks"], ls["observations"], ls["parameters"], ls["verbatim_reference"]["s"], ls["verbatim_reference"]["nfev"]),
                "ekf_update_frames_per_s_by_rays": ekf}
    except Exception:
        return None


def solver_phase_times(prob, x0, ref_pose, n_iter=3):
    """CUDA-event time of every phase of an LM iteration: the library prints them to stderr when PTZBA_TRACE is set (read at
    every solver call), so a few extra iterations run with fd 2 pointed at a temporary file.  Returns {phase: median us}."""
    import re
    import tempfile
    sys.stderr.flush()
    saved = os.dup(2)
    tmp = tempfile.TemporaryFile(mode="w+b")
    os.environ["PTZBA_TRACE"] = "1"
    try:
        os.dup2(tmp.fileno(), 2)
        for _ in range(n_iter):
            prob.lm_iteration(x0, ref_pose, alpha=1e-3)
    finally:
        os.dup2(saved, 2)
        os.close(saved)
        os.environ.pop("PTZBA_TRACE", None)
    tmp.seek(0)
    text = tmp.read().decode(errors="replace")
    tmp.close()
    phases = {}
    for name, us in re.findall(r"([A-Za-z_+|0-9]+)=([0-9.]+)us", text):
        phases.setdefault(name, []).append(float(us))
    return {k: statistics.median(v) for k, v in phases.items()}


def ekf_flops(n_matched, lu_route):
    """Algorithmic FP64 flops of one EKF update with n matched rays (m = 2n rows, s = 3 + 2n state columns), BASELINE.md section 4:
    Cholesky route m^3/3 + m^2 (s+1) + 2 n^2 m (+ 2 m s); pivoted-LU route 2 m^3/3 + 2 m^2 (s+1) + 2 n^2 m (+ 2 m s)."""
    n = np.asarray(n_matched, dtype=np.float64)
    m, s_ = 2 * n, 3 + 2 * n
    chol = m ** 3 / 3 + m * m * (s_ + 1) + 2 * n * n * m + 2 * m * s_
    lu = 2 * m ** 3 / 3 + 2 * m * m * (s_ + 1) + 2 * n * n * m + 2 * m * s_
    return np.where(np.asarray(lu_route) != 0, lu, chol)


def bench_ekf(ctx, n_seq, n_rays, n_frames, seed0=2000):
    """Batched independent EKF sequences (BASELINE config 4 recipe at a bounded batch): predict+update per frame."""
    import torch
    from ptz_slam_b200 import synth, _lib
    from ptz_slam_b200.ptz_slam import BatchedEkfTracker
    seqs = [synth.make_ekf_sequence(n_rays, n_frames + 1, seed=seed0 + i) for i in range(n_seq)]
    max_obs = max(len(i) for q in seqs for i in q.obs_idx)
    trk = BatchedEkfTracker(np.stack([q.rays0 for q in seqs]), np.stack([q.ptz_gt[0] for q in seqs]), synth.PP_U, synth.PP_V,
                            max_obs, synth.IMAGE_H, synth.IMAGE_W, jacobian_mode=_lib.JAC_CENTRAL_FD, ctx=ctx)
    packed = [trk.pack_observations([q.obs_xy[k] for q in seqs], [q.obs_idx[k] for q in seqs]) for k in range(1, n_frames + 1)]
    # untimed: the first half of the frames.  They take every sequence from its diagonal initial covariance (Cholesky route) to the
    # state a tracker is in for the rest of a 100-frame run: the reference's covariance write-back has made P indefinite and the
    # innovation covariance needs the pivoted-LU route (ptz_slam.py:281-289)
    n_warm = max(1, n_frames // 2)
    for k in range(n_warm):
        trk.step(*packed[k])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tot = 0
    flops = 0.0
    for k in range(n_warm, n_frames):
        m = trk.step(*packed[k])
        tot += int(m.sum())
        flops += float(ekf_flops(m, trk.route()).sum())
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    n_lu = int(trk.route().sum())
    trk.close()
    frames = n_frames - n_warm
    pk = fp64_peak()
    ach = flops / dt / 1e12
    return {"workload": "%d independent sequences x %d rays, frames %d..%d of every sequence timed, host observations in / matched counts out" %
                        (n_seq, n_rays, n_warm + 1, n_frames),
            "sequence_frames_per_s": n_seq * frames / dt, "matched_obs_per_s": tot / dt,
            "mean_matched_rays": tot / (n_seq * frames), "ms_per_frame_batch": 1e3 * dt / frames,
            "sequences_on_lu_route_at_end": n_lu,
            "roofline": {"bound": "fp64", "achieved": ach, "peak": pk["dfma_tflops"], "unit": "TFLOP/s", "frac": ach / pk["dfma_tflops"],
                         "peak_kind": pk["source"], "peak_dmma": pk["dmma_tflops"],
                         "flops": "algorithmic: factorisation of S + triangular solves of [G | y] + covariance down-date, per route (bench.py:ekf_flops)"}}


def bench_projection(ctx, stream, n_cam=4096, n_ray=2000, reps=50):
    """North-star subsystem 1 at the BASELINE config 4 shape (one pose per sequence x all rays): k_project_grid on device-resident
    buffers, CUDA events on the library's stream, rotating over output buffers larger than the L2 in total.  Algorithmic bytes:
    16 B per ray read + 16 B per (camera, ray) pixel written.  Same measurement as scripts/proj_bench.py."""
    import torch
    from ptz_slam_b200 import _lib, synth
    rng = np.random.default_rng(1)
    ptz = torch.from_numpy(np.stack([rng.uniform(50, 70, n_cam), rng.uniform(-10, -6, n_cam), rng.uniform(2000, 4000, n_cam)], 1)).cuda()
    rays = torch.from_numpy(synth.make_ray_cloud(n_ray, 3)).cuda()
    n_buf = max(1, int(np.ceil(300e6 / (n_cam * n_ray * 16))))
    outs = [torch.empty(n_cam * n_ray * 2, dtype=torch.float64, device="cuda") for _ in range(n_buf)]
    P = lambda t: _lib.ptr(int(t.data_ptr()))

    def proj(i):
        ctx.check(ctx.lib.ptzba_project(ctx.handle, _lib.DEVICE, n_cam, P(ptz), synth.PP_U, synth.PP_V, None, n_ray, P(rays), P(outs[i % n_buf])))
    for i in range(5):
        proj(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(reps):
        proj(i)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    nbytes = n_ray * 16 + n_cam * n_ray * 16
    peaks, peak_kind = measured_peaks()
    gbs = nbytes / ms / 1e6
    return {"workload": "%d poses x %d rays (BASELINE config 4 shape), device-resident, %d output buffers in rotation" % (n_cam, n_ray, n_buf),
            "pairs_per_s": n_cam * n_ray / (ms * 1e-3), "us_per_launch": ms * 1e3,
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                         "peak_kind": peak_kind, "kernel": "k_project_grid", "algorithmic_bytes": nbytes}}


def bench_cfg2(ctx, n_rays=3000, n_frames=20):
    """BASELINE config 2 shape: ONE sequence over the 3 000-ray soccer cloud through the drop-in PtzSlam (filter state resident on the
    GPU; per frame the observations go in and the pose comes back).  The first 20 frames of the 300: the reference's recursion
    itself diverges on noisy synthetic data after ~25 frames (its covariance write-back), later frames would
    time a filter that has lost track.  cpu_baseline: the numpy restatement of the reference's ekf_update (oracle) on 3 frames."""
    import torch
    from ptz_slam_b200 import synth
    from ptz_slam_b200.ptz_camera import PTZCamera
    from ptz_slam_b200.ptz_slam import PtzSlam
    seq = synth.make_ekf_sequence(n_rays, n_frames + 2, seed=1002, keep_prob=1.0)
    cam = PTZCamera((synth.PP_U, synth.PP_V), np.zeros(3), np.eye(3))
    cam.set_ptz(seq.ptz_gt[0])
    slam = PtzSlam()
    slam.init_rays(seq.rays0, cam)
    slam.predict()
    slam.ekf_update(seq.obs_xy[1], seq.obs_idx[1], synth.IMAGE_H, synth.IMAGE_W)      # warm-up frame (uploads the state once)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tot = 0
    for k in range(2, n_frames + 2):
        slam.predict()
        tot += slam.ekf_update(seq.obs_xy[k], seq.obs_idx[k], synth.IMAGE_H, synth.IMAGE_W)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    err = np.abs(slam.current_camera.get_ptz() - seq.ptz_gt[n_frames + 1])
    out = {"workload": "1 sequence x %d rays, %d timed frames, PtzSlam.predict + ekf_update on the resident state" % (n_rays, n_frames),
           "frames_per_s": n_frames / dt, "matched_obs_per_s": tot / dt, "mean_matched_rays": tot / n_frames, "ms_per_frame": 1e3 * dt / n_frames,
           "final_pose_error_vs_ground_truth": [float(e) for e in err]}
    from oracle import ptz_oracle as O
    s = O.EkfState(seq.rays0, seq.ptz_gt[0], synth.PP_U, synth.PP_V)
    O.ekf_predict(s)
    O.ekf_update(s, seq.obs_xy[1], seq.obs_idx[1], synth.IMAGE_H, synth.IMAGE_W)
    t0 = time.perf_counter()
    for k in range(2, 5):
        O.ekf_predict(s)
        O.ekf_update(s, seq.obs_xy[k], seq.obs_idx[k], synth.IMAGE_H, synth.IMAGE_W)
    cdt = (time.perf_counter() - t0) / 3
    out["cpu_baseline"] = {"value": 1.0 / cdt, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
                           "sample": "3 frames of the same sequence through oracle.ekf_update (numpy restatement of ptz_slam.py:210-289, BLAS threads)"}
    return out


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def pin_to_gpu_numa_node(local):
    """Multi-rank runs: bind this process (and the pinned host buffers it first-touches afterwards) to the CPUs of the NUMA
    node its GPU hangs off, so that the host<->device copies of the e2e leg do not cross the socket interconnect.  Best
    effort: returns a one-line description of what was done (or why nothing was) for the JSON line."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        dom, bus, dev = (getattr(pr, k, None) for k in ("pci_domain_id", "pci_bus_id", "pci_device_id"))
        if bus is None:
            return "not pinned: torch does not report the PCI address"
        node_path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom or 0, bus, dev or 0)
        node = int(open(node_path).read().strip())
        if node < 0:
            return "not pinned: the platform reports no NUMA node for the GPU"
        cpus = _parse_cpulist(open("/sys/devices/system/node/node%d/cpulist" % node).read())
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if len(use) < 2:
            return "not pinned: %d of this process's %d CPUs are on NUMA node %d" % (len(use), len(allowed), node)
        os.sched_setaffinity(0, use)
        return "pinned to %d CPUs of NUMA node %d" % (len(use), node)
    except Exception as e:       # no sysfs, no permission, ...: the run goes on unpinned
        return "not pinned: %s" % (str(e).splitlines()[0] if str(e) else type(e).__name__)


def clock_ramp(step, seconds, world, sync, device, block=64):
    """Runs step(0), step(1), ... for about `seconds` in blocks of `block`; returns how many ran.  With world > 1 a step
    holds a collective, so every rank has to run the SAME number of steps: rank 0's clock decides after each block and the
    verdict is broadcast (ranks that each watched their own clock left the loop after different counts and dead-locked
    the 4-GPU run of round 2).  `sync` waits for the enqueued steps (torch.cuda.synchronize on the GPU)."""
    import torch
    t_end = time.perf_counter() + seconds
    i = 0
    go = torch.ones(1, dtype=torch.int32, device=device)
    while True:
        for _ in range(block):
            step(i); i += 1
        sync()
        if world > 1:
            import torch.distributed as dist
            go.fill_(1 if time.perf_counter() < t_end else 0)
            dist.broadcast(go, src=0)
            if int(go.item()) == 0:
                return i
        elif time.perf_counter() >= t_end:
            return i


# -------------------------------------------------------------------------------------------------------------------
# GPU arm
# -------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import ptz_slam_b200  # noqa: F401
    from ptz_slam_b200 import _lib, synth
    from ptz_slam_b200 import bundle_adjustment as BA

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: anything native libraries print meanwhile (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local)
    numa = pin_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = _lib.get_context(local)
    # an explicit side stream: the legacy default stream has handle 0, which the C-ABI reads as "use the context's own
    # stream"; events below are recorded on the very stream the kernels are launched on
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)

    fb_full = make_workload(args.workload, 0)            # the same problem on every rank
    u, v = synth.PP_U, synth.PP_V
    ref_pose = fb_full.ptz_init[0]
    x0 = fb_full.x0()
    kf_range = (0, fb_full.n_pose)
    if world > 1:
        from ptz_slam_b200 import dist as pdist
        cam_s, lm_s, xy_s, kf_range = pdist.shard_by_keyframe(fb_full.cam_idx, fb_full.lm_idx, fb_full.obs_xy, fb_full.n_pose, rank, world)
        fb = synth.FlatBA(cam_s, lm_s, xy_s, fb_full.ptz_gt, fb_full.rays_gt, fb_full.ptz_init, fb_full.rays_init)
    else:
        fb = fb_full
    lm_touched = int(np.unique(fb.lm_idx).size)
    abytes = algorithmic_bytes(fb.n_obs, lm_touched, kf_range[1] - kf_range[0])      # this rank's share of the pass
    R = args.replicas
    if R <= 0:      # enough replicas of the rank's arrays that a pass never finds them in the 126 MB L2; the SAME number on every
        # rank (the shards differ slightly in size, and setting up a replica's exchange is a collective)
        R = max(4, int(np.ceil(2.2 * 126e6 * world / algorithmic_bytes(fb_full.n_obs, fb_full.n_landmark, fb_full.n_pose))))
    probs = [BA.BAProblem(fb.n_pose, fb.n_landmark, fb.cam_idx, fb.lm_idx, fb.obs_xy, u, v, ctx=ctx) for _ in range(R)]
    for p_ in probs:      # kernel tuning knobs (include/ptzba.h PTZBA_OPT_*); the defaults are the measured best
        if args.lm_share is not None:
            p_.set_option(_lib.OPT_FUSED_LM_SHARE, args.lm_share)
    x_dev = [torch.from_numpy(x0).cuda() for _ in range(R)]
    r_dev = [torch.empty(2 * fb.n_obs, dtype=torch.float64, device="cuda") for _ in range(R)]
    comm = None
    n_shared = 0
    if world > 1:
        comm = pdist.Communicator(ctx, rank, world)
        n_shared = [comm.setup_exchange(p) for p in probs][0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step(i):
        p = probs[i % R]
        p.normal_equations_device(x_dev[i % R].data_ptr(), ref_pose, r_dev[i % R].data_ptr())
        if comm is not None:
            comm.allreduce_landmark_blocks(p)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # clock ramp (untimed, in addition to the W warm-up steps): ~0.3 s of passes; every rank runs the same number of them
    clock_ramp(device_step, args.ramp, world, torch.cuda.synchronize, "cuda")
    for i in range(args.warmup):
        device_step(i)
    barrier()
    launches0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for i in range(args.steps):
        device_step(i)
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count() - launches0
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    total_obs = fb_full.n_obs
    value = total_obs * args.steps / (ms * 1e-3)

    # ---- parity of the sharded pass: every rank against a single-GPU pass over the WHOLE problem ----------------------
    parity = None
    if world > 1:
        import ctypes as _ct
        N, M = fb.n_pose, fb.n_landmark
        device_step(0)
        U, gc, V, gl = np.empty((N, 6)), np.empty((N, 3)), np.empty((M, 3)), np.empty((M, 2))
        cst = _ct.c_double()
        ctx.check(ctx.lib.ptzba_ba_get_blocks(probs[0].handle, _lib.ptr(U), _lib.ptr(gc), _lib.ptr(V), _lib.ptr(gl), _ct.byref(cst)))
        whole = BA.BAProblem(fb_full.n_pose, fb_full.n_landmark, fb_full.cam_idx, fb_full.lm_idx, fb_full.obs_xy, u, v, ctx=ctx)
        xw = torch.from_numpy(x0).cuda()
        whole.normal_equations_device(xw.data_ptr(), ref_pose)
        U1, gc1, V1, gl1 = np.empty((N, 6)), np.empty((N, 3)), np.empty((M, 3)), np.empty((M, 2))
        c1 = _ct.c_double()
        ctx.check(ctx.lib.ptzba_ba_get_blocks(whole.handle, _lib.ptr(U1), _lib.ptr(gc1), _lib.ptr(V1), _lib.ptr(gl1), _ct.byref(c1)))
        whole.close()
        own = slice(max(kf_range[0], 1), kf_range[1])
        mine = np.unique(fb.lm_idx)
        rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)) if a.size else 0.0
        d = [rel(U[own], U1[own]), rel(gc[own], gc1[own]), rel(V[mine], V1[mine]), rel(gl[mine], gl1[mine]),
             abs(cst.value - c1.value) / c1.value]
        t = torch.tensor(d, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        d = [float(q) for q in t.tolist()]
        parity = {"vs": "single-GPU fused pass over the whole problem, same x", "max_rel_diff": {"U": d[0], "g_c": d[1], "V": d[2], "g_l": d[3], "cost": d[4]},
                  "tolerance": 1e-9, "checked": "every rank: U/g_c of its keyframes, V/g_l of the landmarks it observes, the cost"}
        assert max(d) < 1e-9, parity

    # ---- kernel-only duration of the fused kernel (events bracketing each launch) -> roofline ----------------------
    import ctypes
    ctx.check(ctx.lib.ptzba_profile_begin(ctx.handle))
    for i in range(args.steps):
        probs[i % R].normal_equations_device(x_dev[i % R].data_ptr(), ref_pose, r_dev[i % R].data_ptr())
    n_l, tot = ctypes.c_int32(0), ctypes.c_double(0.0)
    ctx.check(ctx.lib.ptzba_profile_end(ctx.handle, ctypes.byref(n_l), ctypes.byref(tot)))
    kernel_ms = tot.value / max(n_l.value, 1)
    peaks, peak_kind = measured_peaks()
    achieved = abytes / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "fused_traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            if tj.get("workload") == args.workload:
                traffic = tj.get("dram_bytes_per_launch")
        except Exception:
            pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_kind": peak_kind,
                "kernel": "fused pass (k_ba_fused: the landmark-major and the keyframe-major CTA role in one launch)", "kernel_ms": kernel_ms, "algorithmic_bytes": abytes}

    # ---- e2e: HOST (pinned) buffers through the C-ABI, copies inside the timed region ------------------------------
    # The call a user of the normal equations makes: x in; the U/g_c blocks of the rank's keyframes, the V/g_l blocks of the
    # landmarks it observes and the cost out (the residual VECTOR is not an input of any solver step: it is opt-in and timed
    # separately below).  ptzba_ba_normal_equations_begin / ptzba_ba_wait pipeline the copies of one problem replica against the
    # kernels of the next (up to 3 passes in flight); every step still moves its own x in and its own blocks out.
    import ctypes as _ct2
    N, M = fb.n_pose, fb.n_landmark
    lm_lo, lm_hi = (int(fb.lm_idx.min()), int(fb.lm_idx.max()) + 1) if fb.n_obs else (0, 0)
    lm_rule = "the span of the landmark ids the rank observes"
    if world > 1:      # every landmark is downloaded by ONE rank: the one that owns its first keyframe (a contiguous id range)
        own_r = pdist.owned_landmark_range(fb_full.cam_idx, fb_full.lm_idx, fb_full.n_landmark, kf_range)
        if own_r is not None:
            lm_lo, lm_hi = own_r
            lm_rule = "the landmarks the rank owns (first observed by one of its keyframes: ids %d..%d)" % (lm_lo, lm_hi)
    nk, nl = kf_range[1] - kf_range[0], lm_hi - lm_lo
    pin = lambda n: torch.empty(max(n, 1), dtype=torch.float64).pin_memory().numpy()
    hx = [pin(len(x0)) for _ in range(R)]
    for h in hx:
        h[:] = x0
    hU, hgc, hV, hgl = [pin(nk * 6) for _ in range(R)], [pin(nk * 3) for _ in range(R)], [pin(nl * 3) for _ in range(R)], [pin(nl * 2) for _ in range(R)]
    costs = [_ct2.c_double(0.0) for _ in range(R)]
    e2e_steps = max(3, min(args.steps, 100))

    depth = min(3, R)          # passes in flight: H2D of one, kernels of the next, D2H of the one before

    def e2e_loop(n):
        for i in range(n):
            r_ = i % R
            if i >= depth:
                probs[(i - depth) % R].wait()
            probs[r_].normal_equations_begin(hx[r_], ref_pose, kf_range, (lm_lo, lm_hi), hU[r_], hgc[r_], hV[r_], hgl[r_], costs[r_])
        for i in range(max(0, n - depth), n):
            probs[i % R].wait()

    e2e_loop(min(args.warmup, 4))
    barrier()
    t0 = time.perf_counter()
    e2e_loop(e2e_steps)
    e2e_s = time.perf_counter() - t0
    # the device-resident pass of the same problem must give the same blocks
    chkU = np.empty((N, 6)); chkV = np.empty((M, 3)); chkc = _ct2.c_double()
    device_step(0)
    ctx.check(ctx.lib.ptzba_ba_get_blocks(probs[0].handle, _lib.ptr(chkU), None, _lib.ptr(chkV), None, _ct2.byref(chkc)))
    e2e_ok = bool(np.allclose(hU[0].reshape(-1, 6)[:nk], chkU[kf_range[0]:kf_range[1]], rtol=1e-12, atol=0) and
                  np.allclose(hV[0].reshape(-1, 3)[:nl], chkV[lm_lo:lm_hi], rtol=1e-12, atol=0) and
                  abs(costs[0].value - chkc.value) <= 1e-12 * chkc.value)
    if world == 1:
        assert e2e_ok, "the blocks of the host-buffer pass differ from the device-resident pass"
    else:       # the owner-based download ranges have not run on hardware yet: a mismatch is reported in the line, on every rank's behalf
        t = torch.tensor([0.0 if e2e_ok else 1.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ok = bool(float(t.item()) == 0.0)
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {"value": total_obs * e2e_steps / e2e_s, "unit": "obs/s", "h2d_bytes_per_step": int(8 * (len(x0) + 3)),
           "d2h_bytes_per_step": int(8 * (nk * 9 + nl * 5 + 1)), "steps": e2e_steps, "blocks_equal_device_resident_pass": e2e_ok,
           "ms_per_step": 1e3 * e2e_s / e2e_steps,
           "call": "ptzba_ba_normal_equations_begin / ptzba_ba_wait (pinned host buffers, copies of one replica overlap the kernels of the "
                   "next): x in; U, g_c of the rank's keyframes, V, g_l of " + lm_rule + ", cost out; bytes are per rank"}
    # the same pass with the residual vector requested as well (synchronous call, 16 B per observation more to download)
    hr = pin(2 * fb.n_obs)
    fU, fgc, fV, fgl = pin(N * 6), pin(N * 3), pin(M * 3), pin(M * 2)
    n_r = max(3, min(args.steps, 20))
    probs[0].normal_equations_into(hx[0], ref_pose, hr, fU, fgc, fV, fgl)
    barrier()
    t0 = time.perf_counter()
    for i in range(n_r):
        probs[i % R].normal_equations_into(hx[0], ref_pose, hr, fU, fgc, fV, fgl)
    torch.cuda.synchronize()
    er_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([er_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        er_s = float(t.item())
    e2e["with_residual_vector"] = {"value": total_obs * n_r / er_s, "unit": "obs/s", "ms_per_step": 1e3 * er_s / n_r,
                                   "d2h_bytes_per_step": int(8 * (2 * fb.n_obs + N * 9 + M * 5 + 1)),
                                   "call": "ptzba_ba_normal_equations(mem=HOST) with the residual vector (synchronous)"}

    # ---- BA LM iterations/s (fused pass + Schur + Cholesky + back-substitution + trial residual pass) --------------
    lm = None
    if world == 1 and not args.no_lm:
        for _ in range(2):
            probs[0].lm_iteration(x0, ref_pose, alpha=1e-3)
        torch.cuda.synchronize()
        n_lm_it = 10
        t0 = time.perf_counter()
        for _ in range(n_lm_it):
            probs[0].lm_iteration(x0, ref_pose, alpha=1e-3)
        torch.cuda.synchronize()
        lm_s = (time.perf_counter() - t0) / n_lm_it
        t0 = time.perf_counter()
        xs, rep = probs[0].solve(x0, ref_pose, ftol=1e-4)
        solve_s = time.perf_counter() - t0
        lm = {"lm_iters_per_s": 1.0 / lm_s, "ms_per_lm_iter": 1e3 * lm_s,
              "solve": {"ftol": 1e-4, "ms": 1e3 * solve_s, "nfev": rep["nfev"], "njev": rep["njev"], "status": rep["status"],
                        "n_factor": rep["n_factor"], "cost0": rep["cost0"], "cost": rep["cost"],
                        "call": "BAProblem.solve -> ptzba_ba_solve(mem=HOST), the reference's entry point (bundle_adjustment.py:200-208): "
                                "x in and x out through host buffers inside the timed region, wall clock",
                        "h2d_bytes": int(8 * len(x0)), "d2h_bytes": int(8 * len(x0))}}

    if lm is not None and not args.no_cpu_baseline:
        # the same iteration on the host cores: C port for the passes, scipy.sparse / LAPACK for Schur + Cholesky (one iteration,
        # ~5-10 s); its predicted reduction and trial cost are printed beside the GPU's (informational: the asserted parity of the
        # solver is in tests/test_gpu_ba.py and tests/test_gpu_configs.py)
        try:
            from oracle import ptz_oracle as O, c_port
            t0 = time.perf_counter()
            _, pred_c, trial_c = O.ba_lm_iteration_sparse(x0, fb.n_pose, ref_pose, fb.cam_idx, fb.lm_idx, fb.obs_xy, u, v, 1e-3,
                                                         fused=c_port.ba_fused, residual=c_port.ba_residual)
            c_s = time.perf_counter() - t0
            pred_g, trial_g = probs[0].lm_iteration(x0, ref_pose, alpha=1e-3)
            lm["lm_cpu_baseline"] = {"value": 1.0 / c_s, "unit": "LM iterations/s", "cores": c_port.max_threads(), "kind": "port",
                                     "sample": "1 iteration of the same problem at the same damping: oracle.ba_lm_iteration_sparse (passes: "
                                               "oracle/ptz_oracle_c.c on all threads; Schur complement: scipy.sparse, one thread; Cholesky: LAPACK)",
                                     "gpu_vs_cpu_rel_diff": {"predicted_reduction": abs(pred_g - pred_c) / abs(pred_c),
                                                             "trial_cost": abs(trial_g - trial_c) / abs(trial_c)}}
        except Exception as e:
            lm["lm_cpu_baseline"] = {"error": (str(e).splitlines() or [type(e).__name__])[0]}

    if lm is not None:
        # FP64-pipe roofline of the dense step of the solve: the Cholesky of the reduced camera system
        try:
            ph = solver_phase_times(probs[0], x0, ref_pose)
            n_red = 3 * (fb.n_pose - 1)
            t_potrf = ph["potrf+block_inverses"] * 1e-6
            pk = fp64_peak()
            ach = n_red ** 3 / 3.0 / t_potrf / 1e12
            lm["roofline_dense"] = {"bound": "fp64", "achieved": ach, "peak": pk["dfma_tflops"], "unit": "TFLOP/s", "frac": ach / pk["dfma_tflops"],
                                    "peak_kind": pk["source"], "kernel_us": ph["potrf+block_inverses"],
                                    "kernel": "k_potrf_coop: Cholesky of the reduced camera system, order %d (n^3/3 flops), + the 128x128 block "
                                              "inverses, one persistent cooperative launch" % n_red,
                                    "note": "latency-bound at this order (panels x device-wide barriers), not FP64-bound: DESIGN.md section 6"}
            lm["lm_phases_us"] = ph
        except Exception as e:
            lm["roofline_dense"] = {"error": (str(e).splitlines() or [type(e).__name__])[0]}

    if world > 1 and not args.no_lm:
        # distributed solve, STRONG scaling of one problem of the named workload: every rank holds the whole observation
        # list (replicated data) and visits its landmark / keyframe-major slices (partitioned work); partial blocks, the
        # reduced camera system, right-hand sides and landmark steps are all-reduced inside the library
        from ptz_slam_b200 import dist as pdist
        g = fb_full
        gp = BA.BAProblem(g.n_pose, g.n_landmark, g.cam_idx, g.lm_idx, g.obs_xy, u, v, ctx=ctx)
        lm_range, cm_range = pdist.solve_partition(g.lm_idx, g.n_landmark, world)[rank]
        gp.set_partition(rank, world, lm_range, cm_range)
        gx0, gref = g.x0(), g.ptz_init[0]
        for _ in range(2):
            gp.lm_iteration(gx0, gref, alpha=1e-3)
        barrier()
        n_lm_it = 10
        t0 = time.perf_counter()
        for _ in range(n_lm_it):
            gp.lm_iteration(gx0, gref, alpha=1e-3)
        torch.cuda.synchronize()
        t = torch.tensor([(time.perf_counter() - t0) / n_lm_it], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        lm_s = float(t.item())
        t0 = time.perf_counter()
        xs, rep = gp.solve(gx0, gref, ftol=1e-4)
        solve_s = time.perf_counter() - t0
        gp.close()
        # parity of the distributed solve: the same problem solved on this rank's GPU alone
        one = BA.BAProblem(g.n_pose, g.n_landmark, g.cam_idx, g.lm_idx, g.obs_xy, u, v, ctx=ctx)
        x1, rep1 = one.solve(gx0, gref, ftol=1e-4)
        one.close()
        nc = 3 * (g.n_pose - 1)
        dx = np.abs(xs - x1)
        ang = float(np.radians(max(dx[0:nc:3].max(), dx[1:nc:3].max(), dx[nc:].max())))
        foc = float(dx[2:nc:3].max())
        t = torch.tensor([ang, foc, float(np.abs(xs).sum())], dtype=torch.float64, device="cuda")
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tmin = t.clone(); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        solve_parity = {"vs": "single-GPU solve of the same problem (ftol 1e-4)", "max_angle_diff_rad": float(tmax[0]), "max_focal_diff_px": float(tmax[1]),
                        "status": [rep["status"], rep1["status"]], "nfev": [rep["nfev"], rep1["nfev"]],
                        "x_identical_on_all_ranks": bool(float(tmax[2]) == float(tmin[2])),
                        "tolerance": "1e-6 rad, 1e-3 px (BASELINE.json)"}
        assert solve_parity["max_angle_diff_rad"] < 1e-6 and solve_parity["max_focal_diff_px"] < 1e-3 and solve_parity["x_identical_on_all_ranks"], solve_parity
        lm = {"lm_iters_per_s": 1.0 / lm_s, "ms_per_lm_iter": 1e3 * lm_s, "lm_scaling": "strong",
              "lm_workload": "%s: ONE problem of %d keyframes x %d landmarks x %d observations solved by %d ranks" %
                             (args.workload, g.n_pose, g.n_landmark, g.n_obs, world),
              "solve": {"ftol": 1e-4, "ms": 1e3 * solve_s, "nfev": rep["nfev"], "njev": rep["njev"], "status": rep["status"],
                        "n_factor": rep["n_factor"], "cost0": rep["cost0"], "cost": rep["cost"]},
              "solve_parity": solve_parity}

    # ---- batched EKF tracking (config 4 shape, bounded batch): sequence-frames/s and matched observations/s ----------
    ekf = None
    if not args.no_ekf:
        # independent sequences shard over the ranks with no collective; rates add up over the ranks
        ekf = bench_ekf(ctx, max(1, args.ekf_seqs), args.ekf_rays, args.ekf_frames, seed0=2000 + rank * args.ekf_seqs)
        if world > 1:
            t = torch.tensor([ekf["ms_per_frame_batch"], ekf["matched_obs_per_s"] * ekf["ms_per_frame_batch"] * 1e-3],
                             dtype=torch.float64, device="cuda")          # (ms per frame batch, matched observations per frame batch)
            tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            ms_ekf = float(tmax[0].item())
            ekf["workload"] = "%d ranks x (%s), sharded with no collective" % (world, ekf["workload"])
            ekf["sequence_frames_per_s"] = world * args.ekf_seqs / (ms_ekf * 1e-3)
            ekf["matched_obs_per_s"] = float(tsum[1].item()) / (ms_ekf * 1e-3)
            ekf["ms_per_frame_batch"] = ms_ekf

    cfg2 = None
    if rank == 0 and not args.no_ekf:
        cfg2 = bench_cfg2(ctx)

    projection = None
    if rank == 0 and world == 1:
        try:
            projection = bench_projection(ctx, stream)
        except Exception as e:
            projection = {"error": (str(e).splitlines() or [type(e).__name__])[0]}

    clocks = sampler.stop() if rank == 0 else None

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        c = time_cpu_port(fb, 20, 1, budget_s=15.0)
        cpu = {"value": c["value"], "unit": "obs/s", "cores": c["cores"], "kind": "port",
               "sample": "%d full fused passes over all %d observations of %s (oracle/ptz_oracle_c.c, %d threads)" %
                         (c["steps"], fb.n_obs, args.workload, c["cores"])}

    if rank == 0:
        line = {
            "metric": "ba_residual_jacobian_normal_eq_obs_per_s", "value": value, "unit": "obs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s keyframe BA: %d keyframes x %d ray landmarks x %d observations%s" %
                       (args.workload, fb_full.n_pose, fb_full.n_landmark, fb_full.n_obs,
                        " per GPU" if world == 1 else ", ONE problem sharded by keyframe over %d GPUs" % world),
                       "l2": "rotating over %d replicas of the rank's observation arrays (%.0f MB per pass and rank: more than the 126 MB L2 in total)" %
                             (R, abytes / 1e6),
                       "parallelism": ("keyframe-sharded observations (dist.shard_by_keyframe), 1 rank per GPU; every timed pass ends with one ncclAllReduce "
                                       "of the cost and of the V/g_l blocks of the landmarks observed by more than one rank (n_shared = %d of %d landmarks, "
                                       "%.2f MB per rank and pass); U/g_c are complete on the rank that owns the keyframe" %
                                       (n_shared, fb.n_landmark, (1 + 5 * n_shared) * 8 / 1e6)) if world > 1 else "1 GPU",
                       "n_shared": int(n_shared), "numa": numa,
                       "step": "one fused residual+Jacobian+normal-equation pass (k_set_params + k_ba_fused)%s" %
                               (" + exchange (k_pack_shared, ncclAllReduce, k_unpack_shared)" if world > 1 else "")},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        vr = verbatim_reference_record()
        if vr:
            line["verbatim_reference_cpu"] = vr
        if parity:
            line["parity"] = parity
        if lm:
            line.update(lm)
        if ekf:
            line["ekf"] = ekf
        if cfg2:
            line["ekf_cfg2"] = cfg2
        if projection:
            line["projection"] = projection
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    for p in probs:
        p.close()
    if world > 1:
        dist.destroy_process_group()


def start_watchdog(seconds):
    """A run that has not finished after `seconds` is a hang (a collective some rank never entered, a wedged device): say so on
    stderr and end THIS process with exit code 124, so that torchrun tears the other ranks down and the box is handed back
    instead of sitting in a collective until somebody else's time limit expires.  0 disables it."""
    if seconds <= 0:
        return None

    def bark():
        sys.stderr.write("bench.py: watchdog: rank %s still running after %d s - aborting the process (exit code 124)\n" %
                         (os.environ.get("RANK", "0"), seconds))
        sys.stderr.flush()
        os._exit(124)

    t = threading.Timer(seconds, bark)
    t.daemon = True
    t.start()
    return t


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--replicas", type=int, default=0, help="problem replicas the passes rotate over (0 = enough to exceed the L2)")
    ap.add_argument("--lm-share", type=int, default=None, help="tuning: per cent of the fused pass's CTA budget for the landmark-major role")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-lm", action="store_true")
    ap.add_argument("--no-ekf", action="store_true")
    ap.add_argument("--ekf-seqs", type=int, default=64)
    ap.add_argument("--ekf-rays", type=int, default=2000)
    ap.add_argument("--ekf-frames", type=int, default=8)
    ap.add_argument("--ramp", type=float, default=0.3, help="seconds of untimed passes before warm-up (clock ramp)")
    ap.add_argument("--watchdog", type=float, default=1500.0, help="abort the process after this many seconds (0 = never); a hang must not hold the box")
    args = ap.parse_args()
    start_watchdog(args.watchdog)
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
