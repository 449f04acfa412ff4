#!/usr/bin/env python
"""Benchmark of the batched MSCKF filter-step (BASELINE.json metric: feature-updates/s and filter-steps/s, % roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--filters B] [--mode qr|gram] [--impl reference]

Workload (config.workload): BASELINE.json configs[3] shape per GPU -- 4096 concurrent filters x 256 features x 30-pose
window (d = 195), FP64, 9 IMU samples per camera frame.  Every filter starts from a steady-state snapshot (the window is
full) taken from a real filter run at that shape (tests/golden/baseline_256x30_updates.npz: P, window poses and tracks as the
reference held them), replicated over the batch with per-filter pixel noise and IMU-noise level (index mod 3).
One step = one filter-step of every filter: restore the snapshot state (device copies, inside the timed region),
9 fused IMU propagations, one update (selection, triangulation, Jacobians, nullspace, gating, compression, Joseph update,
state correction).  Multi-GPU: filters are independent, each rank owns its own 4096 (weak scaling), no data-path
collective; a final gather of the per-filter positions happens outside the timed region.

`--impl reference` times the CPU path (the NumPy oracle, a line-by-line port of the reference's update; the reference
itself is Python and is not present on the GPU box) on all host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FIXTURE = os.path.join(ROOT, "tests", "golden", "baseline_256x30_updates.npz")
N_IMU = 9
FRAME_DT = 0.05
NOISE = [(1e-3, 1e-4, 1e-4, 1e-5), (5e-3, 5e-4, 5e-4, 5e-5), (1e-2, 1e-3, 1e-3, 1e-4)]   # (sa, sg, sba, sbg) low/mid/high


# ----------------------------------------------------------------------------------------------------------------------
# workload
# ----------------------------------------------------------------------------------------------------------------------
def stress_case():
    """BASELINE.json configs[4] shape (1024 features x 60 observations, 60-pose window, d = 375): no reference run exists at
    this size (its update() would need 115 GB), so the snapshot is the self-consistent synthetic case the parity tests use."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import make_case
    return make_case(1024, 60, 60, seed=7)


def load_base_case(n_cams=30, config=4):
    if config == 5:
        return stress_case()
    z = np.load(FIXTURE)
    n = int(z["n_cases"])
    best = None
    for i in range(n):
        c = {k[len(f"c{i}."):]: (z[k].item() if z[k].ndim == 0 else z[k]) for k in z.files if k.startswith(f"c{i}.")}
        if int(c["n_cams"]) == n_cams:
            best = c
    assert best is not None, f"fixture holds no {n_cams}-pose case"
    return best


def filter_inputs(base, index):
    """Per-filter variation of the snapshot: pixel noise on every observation + its IMU-noise level and samples."""
    rng = np.random.default_rng(42 + index)
    uv = base["obs_uv"] + 0.1 * rng.standard_normal(base["obs_uv"].shape)
    sa, sg, sba, sbg = NOISE[index % 3]
    dt = FRAME_DT / N_IMU
    R = np.asarray(base["imu_R"], float)
    t0 = 1.0
    samples = np.zeros((N_IMU, 7))
    samples[:, 0] = t0 + dt * np.arange(1, N_IMU + 1)
    samples[:, 1:4] = 0.05 * np.sin(np.arange(N_IMU)[:, None] * 0.3 + np.arange(3)) + sg * rng.standard_normal((N_IMU, 3))
    samples[:, 4:7] = (R.T @ np.array([0.0, 0.0, -9.81]))[None, :] + 0.2 * np.cos(np.arange(N_IMU)[:, None] * 0.2 + np.arange(3)) \
        + sa * rng.standard_normal((N_IMU, 3))
    qc = np.array([sg ** 2, sbg ** 2, sa ** 2, sba ** 2])
    return uv, samples, qc, t0


def captured_traffic(kernel, filters_per_launch):
    """dram__bytes_read.sum + dram__bytes_write.sum of `kernel` from the committed `ncu --set full` capture
    (profiles/roofline_traffic.json: bytes of one launch over `filters` filters of the benchmark shape), scaled to the
    number of filters one launch of this run covers.  None when there is no capture of that kernel."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get(kernel)
        return float(t["bytes"]) * filters_per_launch / float(t["filters"])
    except Exception:
        return None


def algorithmic_bytes_per_filter_step(F, M, N, n):
    """F (52 M + 104) + 2 d^2 8 + 2 (96 N + 168) + 2 15 d 8 + 56 n."""
    d = 15 + 6 * N
    return F * (52 * M + 104) + 2 * d * d * 8 + 2 * (96 * N + 168) + 2 * 15 * d * 8 + 56 * n


def algorithmic_qr_flops_per_filter(rows, N):
    """explicit QR of the stacked rows, 2 m c^2 - 2/3 c^3."""
    c = 6 * N
    return 2.0 * rows * c * c - (2.0 / 3.0) * c ** 3


def algorithmic_asm_flops_per_feature(M):
    """front_asm_kernel: K = H_x P H_x^T + s^2 I from M (M + 1) / 2 pairs of 2x6 . 6x6 . 6x2 blocks (96 multiply-adds each)."""
    return 192.0 * (M * (M + 1) / 2)


def algorithmic_geo_flops_per_feature(M):
    """front_geo_kernel: ~300 flops per observation of geometry (triangulation, residual, Jacobians, basis of range(H_f))."""
    return 300.0 * M


def algorithmic_chol_flops_per_feature(M):
    """front_chol_kernel: Cholesky of the 2M x 2M matrix with the 4 appended rows [r~; Q1^T]: n^3 / 3 + 4 n^2."""
    n = 2 * M
    return n ** 3 / 3.0 + 4.0 * n * n


def algorithmic_front_flops_per_feature(M):
    """Rows 3-8 for one track of M observations (kernel table): assembling K = H_x P H_x^T + s^2 I from
    M (M + 1) / 2 pairs of 2x6 . 6x6 . 6x2 blocks (96 multiply-adds each), the Cholesky of the 2M x 2M matrix with the 4
    appended rows (n^3 / 3 + 4 n^2), and ~300 flops per observation of geometry.  Lower than SURVEY.md's 0.35-0.45 MFLOP
    estimate (which assumed the two-sided projection); the smaller figure is used so the fraction is not flattered."""
    n = 2 * M
    return 192.0 * (M * (M + 1) / 2) + n ** 3 / 3.0 + 4.0 * n * n + 300.0 * M


def algorithmic_gram_flops_per_filter(n_acc, N):
    """Information-form compression: one rank-3 symmetric update of the 6N x 6N matrix per accepted feature."""
    c = 6 * N
    return 2.0 * 3.0 * n_acc * c * (c + 1) / 2.0


def algorithmic_info_update_flops_per_filter(N):
    """update_info.cu: Cholesky panel of P (c^3/3 + 15 c^2), L^T G L (c^3 + c^3/3), Cholesky of B (c^3/3),
    W = L' C^-T (d c^2), P+ = s^2 W W^T lower (d^2 c)."""
    c = 6.0 * N
    d = c + 15.0
    return c ** 3 / 3 + 15 * c * c + c ** 3 + c ** 3 / 3 + c ** 3 / 3 + d * c * c + d * d * c


# ----------------------------------------------------------------------------------------------------------------------
# CPU arm (oracle port on the host cores)
# ----------------------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    index, n_steps, config = args
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    from oracle import msckf_oracle as o
    base = load_base_case(config=config)
    uv, samples, qc, t0 = filter_inputs(base, index)
    c = dict(base)
    c["obs_uv"] = uv
    t_start = time.perf_counter()
    nfeat = 0
    for _ in range(n_steps):
        st = dict(imu_R=c["imu_R"], imu_t=c["imu_t"], imu_v=c["imu_v"], imu_bg=c["imu_bg"], imu_ba=c["imu_ba"],
                  null_R=c["imu_R"], null_t=c["imu_t"], null_v=c["imu_v"], P=c["P"], timestamp=t0,
                  gravity=c["gravity"], Qc_diag=np.repeat(qc, 3))
        for k in range(N_IMU):
            st["sample"] = samples[k]
            out = o.propagate_case(st)
            st.update(imu_R=out["out_imu_R"], imu_t=out["out_imu_t"], imu_v=out["out_imu_v"], P=out["out_P"],
                      null_R=out["out_imu_R"], null_t=out["out_imu_t"], null_v=out["out_imu_v"],
                      timestamp=out["out_timestamp"])
        cc = dict(c)
        cc.update(P=st["P"], imu_R=st["imu_R"], imu_t=st["imu_t"], imu_v=st["imu_v"])
        o.update_case(cc)
        nfeat += len(c["feat_len"])
    return time.perf_counter() - t_start, nfeat


def run_cpu(n_workers, steps_per_worker, config=4):
    """Each of n_workers processes (1 BLAS thread each) runs `steps_per_worker` filter-steps of its own filter."""
    import multiprocessing as mp
    for k in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(n_workers) as pool:
        res = pool.map(_cpu_worker, [(i, steps_per_worker, config) for i in range(n_workers)])
    wall = time.perf_counter() - t0
    busy = max(r[0] for r in res)          # exclude interpreter start-up / imports: the slowest worker's compute time
    feats = sum(r[1] for r in res)
    return dict(feature_updates_per_s=feats / busy, filter_steps_per_s=n_workers * steps_per_worker / busy,
                wall_s=wall, busy_s=busy)


# ----------------------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile(prefix="clocks_", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            rows = [ln.strip().split(", ") for ln in open(self.path) if ln.strip()]
            sm = [float(r[1]) for r in rows if len(r) >= 9]
            if sm:
                out["sm_mhz"] = float(np.median(sm))
                out["sm_max_mhz"] = float(rows[0][2])
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                for i, nme in enumerate(names):
                    if any(len(r) >= 9 and r[5 + i].strip().lower() == "active" for r in rows):
                        out["reasons"].append(nme)
                out["samples"] = len(sm)
            os.unlink(self.path)
        except Exception:
            pass
        return out


_ORIG_AFFINITY = []


def pin_near_gpu(gpu_index):
    """Restrict this rank to the CPU cores `nvidia-smi topo -m` lists as local to its GPU, BEFORE the pinned host buffers
    are allocated (first touch then places them on that NUMA node): with 8 ranks uploading at once, cross-socket copies
    are what limits the end-to-end step.  Best effort: any failure leaves the affinity untouched."""
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        avail = os.sched_getaffinity(0)
        _ORIG_AFFINITY.append(set(avail))
        for ln in out.splitlines():
            parts = ln.split()
            if not parts or parts[0] != f"GPU{gpu_index}":
                continue
            for tok in parts[1:]:
                if tok[0].isdigit() and ("-" in tok or "," in tok):
                    cores = set()
                    for rng in tok.split(","):
                        lo, _, hi = rng.partition("-")
                        cores.update(range(int(lo), int(hi or lo) + 1))
                    cores &= avail
                    if len(cores) >= 2:
                        os.sched_setaffinity(0, cores)
                        return tok
        return None
    except Exception:
        return None


# ----------------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------------
def run_gpu_snapshot(args):
    import torch
    import torch.distributed as dist
    from monocular_visual_inertial_msckf_b200 import BatchedMSCKF, UpdateParams, _lib
    from monocular_visual_inertial_msckf_b200.sharding import gather_filter_results, max_over_ranks, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = pin_near_gpu(local_rank) if world > 1 else None   # host buffers of a rank live on its GPU's NUMA node

    base = load_base_case(args.window, args.config)
    B = args.filters
    F = len(base["feat_len"])
    M = base["obs_uv"].shape[1]
    N = int(base["n_cams"])
    mode = _lib.COMPRESS_QR if args.mode == "qr" else _lib.COMPRESS_GRAM
    gate = _lib.GATE_BOUNDS if args.gate == "bounds" else _lib.GATE_EXACT
    par = UpdateParams.from_case(base, mode, gate)
    flt = BatchedMSCKF(B, F, M, N, par, device=dev)

    # ---- host-side inputs (pinned): what a caller hands over every step ----
    t0h = time.time()
    h_uv = torch.empty((B, F, M, 2), dtype=torch.float64).pin_memory()
    h_samples = torch.empty((B, N_IMU, 7), dtype=torch.float64).pin_memory()
    h_qc = torch.empty((B, 4), dtype=torch.float64)
    uvn, smn, qcn = h_uv.numpy(), h_samples.numpy(), h_qc.numpy()
    t_start = 1.0
    first_filter, _ = shard_range(B * world, rank, world)       # contiguous shard of the Monte-Carlo batch (weak scaling)
    for b in range(B):
        uvn[b], smn[b], qcn[b], t_start = filter_inputs(base, first_filter + b)

    def rep(x, dtype):
        x = np.asarray(x)
        return torch.from_numpy(np.broadcast_to(x, (B,) + x.shape).astype(dtype).copy()).pin_memory()

    h_dir = rep(base["obs_dir"], np.float64)
    h_conf = rep(base["obs_conf"], np.float64)
    h_slot = rep(np.maximum(base["obs_slot"], 0), np.int32)
    h_len = rep(base["feat_len"], np.int32)
    h_lost = rep(base["lost_cnt"], np.int32)
    h_trk = rep(base["tracked_cnt"], np.int32)
    h_anchor = rep(base["idp_anchor_slot"], np.int32)
    h_base = rep(base["idp_base"], np.float64)
    h_m = rep(base["idp_m"], np.float64)
    h_rho = rep(base["idp_rho"], np.float64)
    step_inputs = [(flt.obs_uv, h_uv), (flt.obs_dir, h_dir), (flt.obs_conf, h_conf), (flt.obs_slot, h_slot),
                   (flt.feat_len, h_len), (flt.lost_cnt, h_lost), (flt.tracked_cnt, h_trk),
                   (flt.idp_anchor_slot, h_anchor), (flt.idp_base, h_base), (flt.idp_m, h_m), (flt.idp_rho, h_rho)]
    h2d_bytes = sum(h.numel() * h.element_size() for _, h in step_inputs) + h_samples.numel() * 8
    samples_dev = torch.empty((B, N_IMU, 7), dtype=torch.float64, device=dev)

    # ---- resident snapshot state (restored at the start of every step) ----
    d = 15 + 6 * N
    P0 = torch.zeros_like(flt.P)
    P0[:, :d, :d] = torch.from_numpy(np.asarray(base["P"], float)).to(dev)
    imu0 = torch.from_numpy(np.concatenate([np.asarray(base["imu_R"], float).flatten(), base["imu_t"], base["imu_v"],
                                            base["imu_bg"], base["imu_ba"]])).to(dev).repeat(B, 1)
    camR0 = torch.from_numpy(np.asarray(base["cam_R"], float).reshape(N, 9)).to(dev).repeat(B, 1, 1)
    camt0 = torch.from_numpy(np.asarray(base["cam_t"], float)).to(dev).repeat(B, 1, 1)
    m0 = h_m.to(dev)
    rho0 = h_rho.to(dev)
    time0 = torch.full((B,), t_start, dtype=torch.float64, device=dev)
    flt.n_cams.fill_(N)
    flt.Qc_diag.copy_(h_qc)
    for dst, src in step_inputs:
        dst.copy_(src)
    samples_dev.copy_(h_samples)
    torch.cuda.synchronize()
    setup_s = time.time() - t0h

    def restore():
        flt.P.copy_(P0)
        flt.imu_state.copy_(imu0)
        flt.cam_R.copy_(camR0)
        flt.cam_t.copy_(camt0)
        flt.idp_m.copy_(m0)
        flt.idp_rho.copy_(rho0)
        flt.imu_time.copy_(time0)

    def step_resident():
        restore()
        flt.propagate(samples_dev)
        flt.update()

    h_out_state = torch.empty((B, 21), dtype=torch.float64).pin_memory()
    h_out_camt = torch.empty((B, N, 3), dtype=torch.float64).pin_memory()
    h_out_status = torch.empty((B, F), dtype=torch.uint8).pin_memory()
    h_out_nacc = torch.empty((B,), dtype=torch.int32).pin_memory()
    d2h_bytes = sum(t.numel() * t.element_size() for t in (h_out_state, h_out_camt, h_out_status, h_out_nacc))

    host_in = dict(obs_uv=h_uv, obs_dir=h_dir, obs_conf=h_conf, obs_slot=h_slot, feat_len=h_len, lost_cnt=h_lost,
                   tracked_cnt=h_trk, idp_anchor_slot=h_anchor, idp_base=h_base, idp_m=h_m, idp_rho=h_rho)
    host_out = dict(imu_state=h_out_state, cam_t=h_out_camt, feat_status=h_out_status, n_accepted=h_out_nacc)

    def step_e2e():
        # the caller's host buffers -> device, the step, the results back to the host: ONE C-ABI call
        # (msckf_filter_step_host) that pipelines upload / kernels / download over chunks of filters
        flt.P.copy_(P0)
        flt.imu_state.copy_(imu0)
        flt.cam_R.copy_(camR0)
        flt.cam_t.copy_(camt0)
        flt.imu_time.copy_(time0)
        flt.step_host(host_in, h_samples, host_out, chunk_filters=args.chunk, overlap_previous=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1), device=dev)

    # ---- warm-up, then the timed K steps ----
    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = flt.lib.msckf_launch_count()
    ms_total = timed(step_resident, args.steps)
    launches = flt.lib.msckf_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else {}
    n_acc = int(flt.n_accepted.sum().item())
    n_rej = int(flt.n_rejected.sum().item())
    rows = 2 * int((flt.feat_len * ((flt.feat_status & 0xF) == 1)).sum().item())

    # ---- the same step with the gate's exact early-out (MSCKF_GATE_BOUNDS), reported beside the headline ----
    gate_bounds = None
    if args.gate == "exact" and mode == _lib.COMPRESS_GRAM:
        status_exact = (flt.feat_status & 0xF).clone()
        flt.cpar.gate_mode = _lib.GATE_BOUNDS
        for _ in range(2):
            step_resident()
        ms_b = timed(step_resident, args.steps)
        n_flag = int(((flt.feat_status & _lib.FEAT_GAMMA_BOUND) != 0).sum().item())
        same = bool(torch.equal(flt.feat_status & 0xF, status_exact))
        flt.cpar.gate_mode = _lib.GATE_EXACT
        gate_bounds = dict(value=B * world * F * args.steps / (ms_b * 1e-3), unit="feature-updates/s",
                           ms_per_step=ms_b / args.steps, decided_by_bounds_rank0=n_flag, features_rank0=B * F,
                           identical_decisions=same)
        step_resident()      # leave the exact-mode outputs in place for the statistics below

    # ---- per-kernel durations (CUDA events on the launching stream), one extra pass outside the timed region ----
    def time_stage(fn, reps=3):
        best = 1e30
        for _ in range(reps):
            restore()
            flt.propagate(samples_dev)
            prep()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best

    stage_ms = {}
    prep = lambda: None  # noqa: E731
    stage_ms["front"] = time_stage(flt.front)
    # the two kernels of the split front end, timed by the library's own events on the launching stream
    import ctypes as C
    flt.lib.msckf_set_stage_timing(1)
    best = (1e30, 1e30, 1e30)
    front_launches = 1
    for _ in range(3):
        restore()
        flt.propagate(samples_dev)
        flt.front()
        g_ms, a_ms, c_ms = C.c_double(0), C.c_double(0), C.c_double(0)
        n_sub = flt.lib.msckf_front_stage_ms(C.byref(g_ms), C.byref(a_ms), C.byref(c_ms))
        if n_sub > 0 and g_ms.value + a_ms.value + c_ms.value < sum(best):
            best = (g_ms.value, a_ms.value, c_ms.value)
            front_launches = n_sub
    flt.lib.msckf_set_stage_timing(0)
    split_front = best[0] < 1e29
    if split_front:
        stage_ms["front_geo"], stage_ms["front_asm"], stage_ms["front_chol"] = best
    prep = flt.front
    stage_ms["compress"] = time_stage(flt.compress)
    prep = lambda: (flt.front(), flt.compress())  # noqa: E731
    stage_ms["apply"] = time_stage(flt.apply_update)
    prep = lambda: None  # noqa: E731

    def only_prop():
        flt.propagate(samples_dev)
    restore()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    only_prop()
    e1.record()
    torch.cuda.synchronize()
    stage_ms["propagate"] = e0.elapsed_time(e1)

    # ---- FP64 pipe peak (measured here: MEASURED_PEAKS.json has no FP64 figure) ----
    import ctypes as C
    sink = torch.zeros(8, dtype=torch.float64, device=dev)
    flops = C.c_double(0)
    fp64_peak = 0.0
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        flt.lib.msckf_probe_fp64(148 * 8, 20000, C.c_void_p(sink.data_ptr()), C.byref(flops),
                                 C.c_void_p(torch.cuda.current_stream().cuda_stream))
        e1.record()
        torch.cuda.synchronize()
        fp64_peak = max(fp64_peak, flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)

    # ---- end to end: host buffers in, results out, copies inside the timed region ----
    for _ in range(2):
        step_e2e()
    e2e_steps = max(2, min(args.steps, 5))
    ms_e2e = timed(step_e2e, e2e_steps)

    # ---- end to end with RESIDENT tracks: the tracks stay in HBM between frames, the host sends only
    #      the frame's new observations (msckf_append_observations_batched) and the IMU samples.  Reported beside `e2e`,
    #      which re-uploads the whole tracks every step as the reference's update(features) boundary implies. ----
    status_ref = (flt.feat_status & 0xF).clone()
    newest = (flt.obs_slot.gather(2, (flt.feat_len.clamp(min=1) - 1).long().unsqueeze(-1)).squeeze(-1) == (N - 1)) & (flt.feat_len > 0)
    len_full = flt.feat_len.clone()
    len_before = (flt.feat_len - newest.to(torch.int32)).contiguous()
    trk_full = flt.tracked_cnt.clone()
    trk_before = (flt.tracked_cnt - newest.to(torch.int32)).contiguous()
    lost_full = flt.lost_cnt.clone()
    last = (len_full.clamp(min=1) - 1).long()
    h_action = (newest.to(torch.uint8) * 2).cpu().pin_memory()
    h_new_uv = h_uv.gather(2, last.cpu().view(B, F, 1, 1).expand(B, F, 1, 2)).squeeze(2).contiguous().pin_memory()
    h_new_conf = h_conf.gather(2, last.cpu().unsqueeze(-1)).squeeze(-1).contiguous().pin_memory()
    d_action = torch.empty_like(h_action, device=dev)
    d_new_uv = torch.empty_like(h_new_uv, device=dev)
    d_new_conf = torch.empty_like(h_new_conf, device=dev)
    inc_h2d = sum(t.numel() * t.element_size() for t in (h_action, h_new_uv, h_new_conf, h_samples))

    def step_incremental():
        restore()
        flt.feat_len.copy_(len_before)
        flt.tracked_cnt.copy_(trk_before)
        flt.lost_cnt.copy_(lost_full)
        d_action.copy_(h_action, non_blocking=True)
        d_new_uv.copy_(h_new_uv, non_blocking=True)
        d_new_conf.copy_(h_new_conf, non_blocking=True)
        samples_dev.copy_(h_samples, non_blocking=True)
        flt.propagate(samples_dev)
        flt.append_observations(d_action, d_new_uv, d_new_conf)
        flt.update()
        h_out_state.copy_(flt.imu_state, non_blocking=True)
        h_out_camt.copy_(flt.cam_t, non_blocking=True)
        h_out_status.copy_(flt.feat_status, non_blocking=True)
        h_out_nacc.copy_(flt.n_accepted, non_blocking=True)

    for _ in range(2):
        step_incremental()
    ms_inc = timed(step_incremental, e2e_steps)
    inc_same = float(((flt.feat_status & 0xF) == status_ref).float().mean().item())
    inc_len_ok = bool(torch.equal(flt.feat_len, len_full))
    flt.feat_len.copy_(len_full)
    flt.tracked_cnt.copy_(trk_full)

    # matching front of add_camera_measurements on the resident tracks (not part of `value`):
    # descriptor matcher (tracks x 300 detections x 64-d float32, every track re-detected with descriptor noise), the
    # epipolar / homography walk of the frame's real proposals over the tracks as they were before the append, the append
    frame_front = None
    try:
        F2, DD = 300, 64
        gen = torch.Generator(device=dev).manual_seed(1234 + rank)
        desc_t = torch.nn.functional.normalize(torch.randn((B, F, DD), device=dev, generator=gen), dim=2).contiguous()
        desc_n = torch.nn.functional.normalize(torch.randn((B, F2, DD), device=dev, generator=gen), dim=2)
        desc_n[:, :min(F, F2)] = desc_t[:, :min(F, F2)]
        desc_n = torch.nn.functional.normalize(desc_n + 0.03 * torch.randn((B, F2, DD), device=dev, generator=gen), dim=2).contiguous()
        kp_n = torch.rand((B, F2, 2), dtype=torch.float64, device=dev) * 400
        sc_n = torch.rand((B, F2), dtype=torch.float64, device=dev)
        n_trk = torch.full((B,), F, dtype=torch.int32, device=dev)
        act_v = torch.empty_like(d_action)
        front_ms = {}
        front_out = {}

        def front_stage(name, fn, prep=lambda: None):
            best = 1e9
            for _ in range(3):
                prep()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                front_out[name] = fn()
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            front_ms[name] = best

        def prep_tracks():
            flt.feat_len.copy_(len_before)
            flt.tracked_cnt.copy_(trk_before)
            flt.lost_cnt.copy_(lost_full)
            act_v.copy_(d_action)

        front_stage("match_descriptors", lambda: flt.match_descriptors(desc_t, desc_n, 0.82, n_trk, None, kp_n, sc_n))
        front_stage("verify_matches", lambda: flt.verify_matches(act_v, d_new_uv, 5.0, 5.0), prep_tracks)
        front_stage("append_observations", lambda: flt.append_observations(act_v, d_new_uv, d_new_conf), prep_tracks)
        m12 = front_out["match_descriptors"][0]
        verdict = front_out["verify_matches"]
        n_prop = int((d_action == 2).sum().item())
        frame_front = dict(
            ms=front_ms, ms_total=sum(front_ms.values()),
            matcher=dict(tracks=F, detections=F2, dim=DD, dtype="f32",
                         matched_fraction=float((m12 >= 0).float().mean().item()),
                         tflops=2.0 * F * F2 * DD * B / (front_ms["match_descriptors"] * 1e-3) / 1e12,
                         note="one similarity product per pair (the reference forms D1 D2^T and D2 D1^T)"),
            verify=dict(proposals=n_prop, passed=int((verdict == 1).sum().item()),
                        observations_walked=int(len_before.sum().item()),
                        gbs=(20.0 * float(len_before.sum().item()) + 18.0 * B * F) / (front_ms["verify_matches"] * 1e-3) / 1e9,
                        thresholds=[5.0, 5.0]),
            note="per camera frame and rank, tracks resident; not included in `value` / `e2e`")
        flt.feat_len.copy_(len_full)
        flt.tracked_cnt.copy_(trk_full)
        flt.lost_cnt.copy_(lost_full)
    except Exception as exc:                                  # the matching front must never cost the headline line
        frame_front = dict(error=repr(exc))

    if world > 1:
        # final gather of per-filter results (positions), outside the timed region
        gathered = gather_filter_results(flt.imu_state, B * world, dst=0)
        assert rank != 0 or gathered.shape[0] == B * world
        tot = torch.tensor([n_acc, n_rej, rows], dtype=torch.float64, device=dev)
        dist.all_reduce(tot)
        n_acc, n_rej, rows = [int(x) for x in tot.tolist()]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    sec = ms_total * 1e-3
    Btot = B * world
    feats_per_step = Btot * F
    value = feats_per_step * args.steps / sec
    fsteps = Btot * args.steps / sec
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    step_bytes = algorithmic_bytes_per_filter_step(F, M, N, N_IMU) * B
    rows_per_rank = rows / world
    acc_per_rank = n_acc / world
    stage_flops = dict(front=algorithmic_front_flops_per_feature(M) * B * F)
    if split_front:
        code = flt.feat_status & 0xF
        n_chol = int((((code == 1) | (code == 2)) & ((flt.feat_status & _lib.FEAT_GAMMA_BOUND) == 0) & (flt.dof > 0)).sum().item())
        stage_flops = dict(front_geo=algorithmic_geo_flops_per_feature(M) * B * F,
                           front_asm=algorithmic_asm_flops_per_feature(M) * n_chol,
                           front_chol=algorithmic_chol_flops_per_feature(M) * n_chol)
    if args.mode == "qr":
        stage_flops["compress"] = algorithmic_qr_flops_per_filter(rows_per_rank / B, N) * B
        stage_flops["apply"] = 8.0 * (15 + 6 * N) ** 3 * B
        names = dict(front="front_kernel", front_geo="front_geo_kernel", front_asm="front_asm_kernel", front_chol="front_chol_kernel",
                     compress="tsqr_kernel", apply="apply_kernel")
    else:
        stage_flops["compress"] = algorithmic_gram_flops_per_filter(acc_per_rank / B, N) * B
        stage_flops["apply"] = algorithmic_info_update_flops_per_filter(N) * B
        names = dict(front="front_kernel", front_geo="front_geo_kernel", front_asm="front_asm_kernel", front_chol="front_chol_kernel",
                     compress="gram_ws_kernel", apply="update_smem_kernel")
    dom = max(stage_flops, key=lambda k: stage_ms[k])      # the dominant kernel of the step
    dom_name, dom_ms, dom_flops = names[dom], stage_ms[dom], stage_flops[dom]
    achieved_tf = dom_flops / (dom_ms * 1e-3) / 1e12
    stage_tflops = {names[k]: stage_flops[k] / (stage_ms[k] * 1e-3) / 1e12 for k in stage_flops}
    traffic = captured_traffic(dom_name, B / max(1, front_launches) if dom.startswith("front_") else B)
    # shared-memory view of the assembly kernel (what actually binds it, profiles/): every lower 2x2 block of K reads two
    # 2x6 Jacobian blocks (192 B) and HALF a padded 6x6 block of P (288 B shared by the two features a warp pairs up)
    smem_view = None
    if split_front:
        smem_peak = 128.0 * 148 * float(clocks.get("sm_max_mhz") or 1965.0) * 1e6 / 1e12      # TB/s: 128 B/clk/SM
        smem_bytes = (M * (M + 1) / 2) * 336.0 * B * F
        smem_view = dict(kernel="front_asm_kernel", achieved=smem_bytes / (stage_ms["front_asm"] * 1e-3) / 1e12,
                         peak=smem_peak, unit="TB/s", frac=smem_bytes / (stage_ms["front_asm"] * 1e-3) / 1e12 / smem_peak,
                         bytes_per_feature=(M * (M + 1) / 2) * 336.0,
                         note="algorithmic shared-memory bytes of the pair loop / kernel time vs 128 B/clk/SM x 148 SMs at "
                              "the max SM clock; ncu: 59 % of peak wavefronts over the whole kernel")
    roofline = dict(bound="tensor", pipe="fp64 (DFMA/DMMA share one nominal peak on B200; tcgen05 has no f64 kind)",
                    kernel=dom_name, achieved=achieved_tf, peak=fp64_peak, unit="TFLOP/s",
                    frac=achieved_tf / fp64_peak if fp64_peak else None, traffic=traffic,
                    peak_source="measured in this run by msckf_probe_fp64 (8 independent DMMA.8x8x4 chains per warp; "
                                "DFMA chains give the same 36-37 TFLOP/s, profiles/fp64_probe.txt)",
                    kernel_ms=dom_ms / (front_launches if dom.startswith("front_") else 1), kernel_ms_per_step=dom_ms,
                    flops_per_launch=dom_flops / (front_launches if dom.startswith("front_") else 1),
                    launches_per_step=(front_launches if dom.startswith("front_") else 1),
                    stage_ms=stage_ms, stage_tflops=stage_tflops,
                    smem=smem_view,
                    step_flops=dict(achieved=sum(stage_flops.values()) / (sec / args.steps) / 1e12, unit="TFLOP/s",
                                    flops_per_filter_step=sum(stage_flops.values()) / B),
                    step_hbm=dict(bound="hbm", achieved=step_bytes / (sec / args.steps) / 1e9, peak=hbm_peak, unit="GB/s",
                                  frac=step_bytes / (sec / args.steps) / 1e9 / hbm_peak, peak_source=hbm_src,
                                  bytes_per_filter_step=algorithmic_bytes_per_filter_step(F, M, N, N_IMU)))

    ncpu = os.cpu_count() or 1
    cpu = dict(value=None)
    if not args.no_cpu:
        if _ORIG_AFFINITY:
            os.sched_setaffinity(0, _ORIG_AFFINITY[0])      # the CPU arm uses every host core again
        nw = min(ncpu, 64)
        r = run_cpu(nw, 1, args.config)
        cpu = dict(value=r["feature_updates_per_s"], unit="feature-updates/s", cores=nw, kind="port",
                   filter_steps_per_s=r["filter_steps_per_s"],
                   sample=f"{nw} filters x 1 filter-step (9 propagations + {F}-feature update) of the same workload, "
                          f"one single-threaded process per core, slowest worker {r['busy_s']:.1f} s")

    e2e_sec = ms_e2e * 1e-3
    line = dict(
        metric="msckf_feature_updates_per_s", value=value, unit="feature-updates/s", n_gpus=world, steps=args.steps,
        warmup=max(args.warmup, 3), ms_per_step=ms_total / args.steps, higher_is_better=True, scaling="weak",
        vs_baseline=None, dtype="f64", data="synthetic",
        filter_steps_per_s=fsteps,
        config=dict(workload=f"BASELINE.json {CONFIGS[args.config]['name']}: {B} filters/GPU x {F} features x {N}-pose window "
                             f"(d={15 + 6 * N}), {N_IMU} IMU samples/frame, FP64; steady-state snapshot replicated with "
                             f"per-filter pixel noise; step = 9 propagations + one update", config=args.config, filters_per_gpu=B, features=F, obs_per_feature=M, window=N,
                    compress=args.mode, gate=args.gate, l2="inputs larger than L2 (%.1f GB touched per step)" % (
                        (step_bytes + 2 * flt.P.numel() * 8) / 1e9),
                    accepted_features=n_acc, gated_out=n_rej),
        roofline=roofline, cpu_baseline=cpu, gate_bounds=gate_bounds,
        e2e=dict(value=feats_per_step * e2e_steps / e2e_sec, unit="feature-updates/s", h2d_bytes_per_step=h2d_bytes * world,
                 d2h_bytes_per_step=d2h_bytes * world, ms_per_step=ms_e2e / e2e_steps, steps=e2e_steps,
                 call="msckf_filter_step_host", chunk_filters=args.chunk,
                 filter_steps_per_s=Btot * e2e_steps / e2e_sec),
        e2e_resident_tracks=dict(value=feats_per_step * e2e_steps / (ms_inc * 1e-3), unit="feature-updates/s",
                                 ms_per_step=ms_inc / e2e_steps, h2d_bytes_per_step=inc_h2d * world,
                                 d2h_bytes_per_step=d2h_bytes * world, steps=e2e_steps,
                                 call="msckf_append_observations_batched + msckf_propagate_batched + msckf_update_batched",
                                 gate_decisions_equal_to_full_upload=inc_same, track_lengths_restored=inc_len_ok,
                                 note="tracks resident in HBM between frames: per step the host sends the new observation of "
                                      "every track (25 B) and the IMU samples instead of the whole tracks"),
        frame_front=frame_front,
        gpu_launches=int(launches), clocks=clocks, setup_s=setup_s, cpu_affinity_rank0=numa)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------------------------
# GPU arm, stream workload (default): heterogeneous filters on device-built streams, whole camera frames
# ----------------------------------------------------------------------------------------------------------------------
CONFIGS = {
    2: dict(filters=1, name="configs[1]: single filter, TartanAir-shaped synthetic sequence, 256 tracked features/frame, 30-pose window"),
    3: dict(filters=64, name="configs[2]: 64 concurrent filters (IMU noise low/mid/high x sequences), 30-pose window, 256 features/update"),
    4: dict(filters=4096, name="configs[3]: 4096 Monte-Carlo filter instances per GPU, 30-pose window, 256 features/update"),
    5: dict(filters=128, name="configs[4] stress: 1024 filters (128 per GPU of 8) x 1024 features/update x 60-pose window (state 375)"),
}


def run_gpu_stream(args):
    """One step = one camera frame of every filter, everything resident: the 9 IMU samples since the last frame
    (propagation), state_augmentation, the frame's detections matched against the tracks (matcher, geometric
    verification, bookkeeping), process_features (get_valid_features + update + correct over every track, deletion of lost
    tracks / unobserved cameras) and, when the window exceeds 30 poses, prune_poorest_camera_states with its own update
    (BatchedMSCKF.frame_step_detections = MSCKF.imu_callback x 9 + feature_callback, MSCKF.py:135-158).  Every filter
    runs its own stream (trajectory seed 42 + global index, IMU-noise level index mod 3, 2 % outliers, 3 % drop-outs), built on
    the device and warmed up until the window cycles between 30 and 31 poses."""
    import ctypes as C

    import torch
    import torch.distributed as dist
    from monocular_visual_inertial_msckf_b200 import BatchedMSCKF, UpdateParams, _lib
    from monocular_visual_inertial_msckf_b200.sharding import gather_filter_results, max_over_ranks, shard_range
    from monocular_visual_inertial_msckf_b200.streams import DeviceStreams

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = pin_near_gpu(local_rank) if world > 1 else None

    B, K, W = args.filters, args.steps, max(args.warmup, 3)
    L, WINDOW, N_WARM, DESC = 256, 30, 34, 32
    F, M, N = L + 32, 32, WINDOW + 1
    mode = _lib.COMPRESS_QR if args.mode == "qr" else _lib.COMPRESS_GRAM
    gate = _lib.GATE_BOUNDS if args.gate == "bounds" else _lib.GATE_EXACT
    first_filter, _ = shard_range(B * world, rank, world)
    t0h = time.time()
    n_frames = N_WARM + W + K + 1
    st = DeviceStreams(range(first_filter, first_filter + B), n_frames, n_landmarks=L, imu_per_frame=N_IMU, frame_dt=FRAME_DT,
                       device=dev, sigma_px=0.1, p_drop=0.03, p_outlier=0.02, desc_dim=DESC, seed=42)
    par = UpdateParams(K=st.K, width=st.width, height=st.height, sigma_image=0.1, min_lost=2, min_tracked=3, use_parallax=True,
                       min_parallax=0.0, compress_mode=mode, gate_mode=gate)
    flt = BatchedMSCKF(B, F, M, N, par, device=dev)
    flt.P.zero_()
    flt.n_cams.zero_()
    flt.feat_len.zero_()
    flt.Qc_diag.copy_(st.Qc_diag)

    frame_kw = dict(min_cosine_similarity=0.82, epipolar_rejection_threshold=1e9, homography_rejection_threshold=1e9)
    graphs = {}

    def frame_eager(inp):
        smp, desc, kp, sc, cnt = inp
        flt.frame_step_detections(smp, st.R_I_C, st.t_I_C, desc, kp, sc, cnt, WINDOW, **frame_kw)

    def frame(inp):
        # small batches are launch-bound (a frame is ~60 launches): the frame is captured once into a CUDA graph and replayed
        g = graphs.get(int(flt.cpar.gate_mode))
        if g is None:
            return frame_eager(inp)
        g.step(*inp)

    for f in range(N_WARM):                       # fill the window (untimed)
        frame_eager(st.frame_inputs(f))
    flt.count_feature_updates = True
    flt.feature_updates = torch.zeros((), dtype=torch.int64, device=dev)
    use_graph = args.graph == "on" or (args.graph == "auto" and B <= 256)
    if use_graph:
        smp0, desc0, kp0, sc0, cnt0 = st.frame_inputs(N_WARM)
        graphs[int(flt.cpar.gate_mode)] = flt.capture_frame_graph(smp0, st.R_I_C, st.t_I_C, desc0, kp0, sc0, cnt0, WINDOW, **frame_kw)
    f0 = N_WARM + W
    inputs = [st.frame_inputs(f0 + k) for k in range(K)]
    for w in range(W):                            # warm-up steps
        frame(st.frame_inputs(N_WARM + w))
    flt.check_window_overflow()
    snap = flt.snapshot()
    torch.cuda.synchronize()
    setup_s = time.time() - t0h
    n_cams0 = int(flt.n_cams.max().item())
    len_stats = flt.feat_len[flt.feat_len > 0].double()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # a rank whose state + tracks fit the 126 MB L2 (configs 2 and 3) gets the L2 flushed between timed steps: every step is
    # timed by its own event pair, a 512 MB buffer is written in between (outside the event pairs)
    resident_bytes = flt.P.numel() * 8 + flt.obs_uv.numel() * 8 * 3.5 + (flt.obs_desc.numel() * 4 if flt.obs_desc is not None else 0)
    flush_l2 = resident_bytes < 512e6
    flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device=dev) if flush_l2 else None

    def timed(fn, steps):
        barrier()
        if not flush_l2:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for k in range(steps):
                fn(k)
            e1.record()
            barrier()
            return max_over_ranks(e0.elapsed_time(e1), device=dev)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for k in range(steps):
            flush_buf.fill_(k & 0xFF)
            evs[k][0].record()
            fn(k)
            evs[k][1].record()
        barrier()
        return max_over_ranks(sum(a.elapsed_time(b) for a, b in evs), device=dev)

    # ---- the timed K steps (inputs resident in HBM) ----
    flt.feature_updates.zero_()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = flt.lib.msckf_launch_count()
    if args.profiler_range:     # `ncu --profile-from-start off`: only the timed region is captured (profiles/*_launches.csv)
        torch.cuda.cudart().cudaProfilerStart()
    ms_total = timed(lambda k: frame(inputs[k]), K)
    if args.profiler_range:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
    launches = flt.lib.msckf_launch_count() - launches0
    if use_graph:                                 # replayed launches do not pass through the library's host-side counter
        launches = K * graphs[int(flt.cpar.gate_mode)].launches_per_replay
    clocks = sampler.stop() if rank == 0 else {}
    fu = int(flt.feature_updates.item())
    state_exact = flt.imu_state.clone()
    flt.check_window_overflow()

    # ---- the same K frames with the gate's exact early-out (identical decisions => identical states) ----
    gate_bounds = None
    if args.gate == "exact" and mode == _lib.COMPRESS_GRAM:
        flt.restore(snap)
        flt.cpar.gate_mode = _lib.GATE_BOUNDS
        if use_graph:
            graphs[int(_lib.GATE_BOUNDS)] = flt.capture_frame_graph(smp0, st.R_I_C, st.t_I_C, desc0, kp0, sc0, cnt0, WINDOW, **frame_kw)
        flt.feature_updates.zero_()
        ms_b = timed(lambda k: frame(inputs[k]), K)
        same = bool(torch.equal(flt.imu_state, state_exact))
        worst = float((flt.imu_state - state_exact).abs().max().item())
        fu_b = int(flt.feature_updates.item())
        flt.cpar.gate_mode = _lib.GATE_EXACT
        gate_bounds = dict(ms_per_step=ms_b / K, identical_states=same, max_state_difference=worst, feature_updates_rank0=fu_b)

    # ---- end to end: every step's inputs come from pinned host memory (uploaded on a copy stream, double-buffered, inside
    #      the timed region), every step's result goes back to the host ----
    Ke = max(2, min(K, 6))
    host_in = [[t.cpu().pin_memory() for t in inputs[k]] for k in range(Ke)]
    dev_in = [[torch.empty_like(t) for t in inputs[0]] for _ in range(2)]
    h2d_bytes = sum(t.numel() * t.element_size() for t in host_in[0])
    h_state = torch.empty((B, 21), dtype=torch.float64).pin_memory()
    h_ncams = torch.empty((B,), dtype=torch.int32).pin_memory()
    d2h_bytes = h_state.numel() * 8 + h_ncams.numel() * 4
    copy_stream = torch.cuda.Stream(device=dev)
    ready = [torch.cuda.Event() for _ in range(2)]
    free = [torch.cuda.Event() for _ in range(2)]

    def upload(k):
        slot = k & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(free[slot])
            for d, h in zip(dev_in[slot], host_in[k]):
                d.copy_(h, non_blocking=True)
            ready[slot].record(copy_stream)

    def step_e2e(k):
        cur = torch.cuda.current_stream(dev)
        if k == 0:
            free[0].record(cur)
            free[1].record(cur)
            upload(0)
        if k + 1 < Ke:
            upload(k + 1)
        cur.wait_event(ready[k & 1])
        frame(dev_in[k & 1])
        free[k & 1].record(cur)
        h_state.copy_(flt.imu_state, non_blocking=True)
        h_ncams.copy_(flt.n_cams, non_blocking=True)

    flt.restore(snap)
    flt.feature_updates.zero_()
    ms_e2e = timed(step_e2e, Ke)
    fu_e2e = int(flt.feature_updates.item())
    flt.count_feature_updates = False

    # ---- per-kernel durations on the state of the first timed frame (one update of every track) ----
    flt.restore(snap)
    smp, desc, kp, sc, cnt = inputs[0]
    flt.propagate(smp)
    flt.augment(st.R_I_C, st.t_I_C)
    n_tracks = (flt.feat_len > 0).sum(1).to(torch.int32)
    m12, m21, action, new_uv, new_conf, _ = flt.match_descriptors(flt.lcm_desc, desc, 0.82, n_tracks, cnt, kp, sc)
    flt.verify_matches(action, new_uv, 1e9, 1e9)
    flt.append_descriptors(action, m12, m21, n_tracks, cnt, desc, new_conf)
    flt.append_observations(action, new_uv, new_conf)
    sub = dict(P=flt.P.clone(), slab=flt.small_slab.clone(), m=flt.idp_m.clone(), rho=flt.idp_rho.clone())
    Nw = int(flt.n_cams.max().item())
    Mw = float(flt.feat_len[flt.feat_len > 0].double().mean().item())

    def sub_restore():
        flt.P.copy_(sub["P"])
        flt.small_slab.copy_(sub["slab"])
        flt.idp_m.copy_(sub["m"])
        flt.idp_rho.copy_(sub["rho"])

    def time_stage(fn, prep=lambda: None):
        best = 1e30
        for _ in range(3):
            sub_restore()
            prep()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best

    stage_ms = dict(front=time_stage(flt.front))
    flt.lib.msckf_set_stage_timing(1)
    best3, front_launches = (1e30, 1e30, 1e30), 1
    for _ in range(3):
        sub_restore()
        flt.front()
        g_ms, a_ms, c_ms = C.c_double(0), C.c_double(0), C.c_double(0)
        n_sub = flt.lib.msckf_front_stage_ms(C.byref(g_ms), C.byref(a_ms), C.byref(c_ms))
        if n_sub > 0 and g_ms.value + a_ms.value + c_ms.value < sum(best3):
            best3, front_launches = (g_ms.value, a_ms.value, c_ms.value), n_sub
    flt.lib.msckf_set_stage_timing(0)
    split_front = best3[0] < 1e29
    if split_front:
        stage_ms["front_geo"], stage_ms["front_asm"], stage_ms["front_chol"] = best3
    stage_ms["compress"] = time_stage(flt.compress, flt.front)
    stage_ms["apply"] = time_stage(flt.apply_update, lambda: (flt.front(), flt.compress()))
    sub_restore()
    flt.front()
    code = flt.feat_status & 0xF
    n_valid = int(((code == 1) | (code == 2)).sum().item())
    n_acc1 = int((code == 1).sum().item())
    lens = flt.feat_len.double()
    valid_mask = ((code == 1) | (code == 2)).double()
    pair_flops = float((192.0 * lens * (lens + 1) / 2 * valid_mask).sum().item())
    chol_flops = float((((2 * lens) ** 3 / 3.0 + 4.0 * (2 * lens) ** 2) * valid_mask).sum().item())
    geo_flops = float((300.0 * lens * (lens > 0)).sum().item())
    flt.restore(snap)
    stage_ms["propagate"] = time_stage(lambda: flt.propagate(smp))
    stage_ms["match_descriptors"] = time_stage(lambda: flt.match_descriptors(flt.lcm_desc, desc, 0.82, n_tracks, cnt, kp, sc))

    # ---- FP64 pipe peak (measured here: MEASURED_PEAKS.json has no FP64 figure) ----
    sink = torch.zeros(8, dtype=torch.float64, device=dev)
    flops = C.c_double(0)
    fp64_peak = 0.0
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        flt.lib.msckf_probe_fp64(148 * 8, 20000, C.c_void_p(sink.data_ptr()), C.byref(flops),
                                 C.c_void_p(torch.cuda.current_stream().cuda_stream))
        e1.record()
        torch.cuda.synchronize()
        fp64_peak = max(fp64_peak, flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)

    tot = torch.tensor([fu, fu_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        gathered = gather_filter_results(flt.imu_state, B * world, dst=0)      # the final gather
        assert rank != 0 or gathered.shape[0] == B * world
        dist.all_reduce(tot)
    fu_all, fu_e2e_all = [float(x) for x in tot.tolist()]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    sec = ms_total * 1e-3
    value = fu_all / sec
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    stage_flops = dict(compress=algorithmic_gram_flops_per_filter(n_acc1 / B, Nw) * B if mode == _lib.COMPRESS_GRAM
                       else algorithmic_qr_flops_per_filter(2.0 * Mw * n_acc1 / B, Nw) * B,
                       apply=(algorithmic_info_update_flops_per_filter(Nw) if mode == _lib.COMPRESS_GRAM else 8.0 * (15 + 6 * Nw) ** 3) * B)
    names = dict(front_geo="front_geo_kernel", front_asm="front_asm_ps_kernel", front_chol="front_chol_kernel",
                 compress=("gram_ws_kernel" if Nw <= 32 else "gram_kernel") if mode == _lib.COMPRESS_GRAM else "tsqr_kernel",
                 apply="update_smem_kernel" if mode == _lib.COMPRESS_GRAM else "apply_kernel")
    if split_front:
        stage_flops.update(front_geo=geo_flops, front_asm=pair_flops, front_chol=chol_flops)
    else:
        stage_flops["front"] = geo_flops + pair_flops + chol_flops
        names["front"] = "front_kernel"
    dom = max(stage_flops, key=lambda k: stage_ms[k])
    dom_ms, dom_flops = stage_ms[dom], stage_flops[dom]
    is_front = dom.startswith("front_")
    traffic = captured_traffic(names[dom], B / max(1, front_launches) if is_front else B)
    update_flops = sum(stage_flops.values())
    step_bytes = algorithmic_bytes_per_filter_step(L, int(round(Mw)), Nw, N_IMU) * B
    roofline = dict(bound="tensor", pipe="fp64 (DFMA/DMMA share one nominal peak on B200; tcgen05 has no f64 kind)",
                    kernel=names[dom], achieved=dom_flops / (dom_ms * 1e-3) / 1e12, peak=fp64_peak, unit="TFLOP/s",
                    frac=dom_flops / (dom_ms * 1e-3) / 1e12 / fp64_peak if fp64_peak else None, traffic=traffic,
                    peak_source="measured in this run by msckf_probe_fp64 (8 independent DMMA.8x8x4 chains per warp; DFMA "
                                "chains give the same 36-37 TFLOP/s, profiles/fp64_probe.txt)",
                    kernel_ms=dom_ms / (front_launches if is_front else 1), kernel_ms_per_update=dom_ms,
                    flops_per_launch=dom_flops / (front_launches if is_front else 1),
                    launches_per_update=(front_launches if is_front else 1),
                    measured_on=f"one update of every track at the first timed frame's state (window {Nw} poses, mean track "
                                f"length {Mw:.1f}, {n_valid} valid / {n_acc1} accepted features on rank 0); a frame runs one such "
                                f"update, plus a second one over the tracks of the two pruned cameras when the window exceeds 30",
                    stage_ms=stage_ms, stage_tflops={names[k]: stage_flops[k] / (stage_ms[k] * 1e-3) / 1e12 for k in stage_flops},
                    update_flops=dict(achieved=update_flops / (sum(stage_ms[k] for k in stage_flops) * 1e-3) / 1e12, unit="TFLOP/s",
                                      flops_per_filter_update=update_flops / B),
                    step_hbm=dict(bound="hbm", achieved=step_bytes / (sec / K) / 1e9, peak=hbm_peak, unit="GB/s",
                                  frac=step_bytes / (sec / K) / 1e9 / hbm_peak, peak_source=hbm_src,
                                  bytes_per_filter_step=algorithmic_bytes_per_filter_step(L, int(round(Mw)), Nw, N_IMU),
                                  note="SURVEY.md section 8d: the path is FP64-bound, not HBM-bound, at these shapes"))

    ncpu = os.cpu_count() or 1
    cpu = dict(value=None)
    if not args.no_cpu:
        if _ORIG_AFFINITY:
            os.sched_setaffinity(0, _ORIG_AFFINITY[0])
        nw = min(ncpu, 64)
        r = run_cpu(nw, 1)
        cpu = dict(value=r["feature_updates_per_s"], unit="feature-updates/s", cores=nw, kind="port",
                   filter_steps_per_s=r["filter_steps_per_s"],
                   sample=f"{nw} filters x 1 filter-step (9 propagations + one 256-feature update at a 30-pose window, the "
                          f"steady-state snapshot captured from the reference at this shape), one single-threaded process per "
                          f"core, slowest worker {r['busy_s']:.1f} s")

    e2e_sec = ms_e2e * 1e-3
    line = dict(
        metric="msckf_feature_updates_per_s", value=value, unit="feature-updates/s", n_gpus=world, steps=K, warmup=W,
        ms_per_step=ms_total / K, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
        filter_steps_per_s=B * world * K / sec, feature_updates_per_filter_step=fu_all / (B * world * K),
        config=dict(workload=f"BASELINE.json {CONFIGS[args.config]['name']}: {B} filters/GPU, each on its own device-built stream "
                             f"(trajectory seed 42 + index, IMU noise low/mid/high by index mod 3, 0.1 px pixel noise, 2 % outliers, "
                             f"3 % drop-outs), {L} landmarks, window cycling {WINDOW}/{WINDOW + 1} poses, {N_IMU} IMU samples/frame, FP64; "
                             f"step = one whole camera frame (propagation, augmentation, matching, update(s), pruning)",
                    config=args.config, filters_per_gpu=B, landmarks=L, track_capacity=F, window=WINDOW, window_at_start=n_cams0,
                    mean_track_length=float(len_stats.mean().item()), min_track_length=float(len_stats.min().item()),
                    compress=args.mode, gate=args.gate, descriptor_dim=DESC,
                    cuda_graph=bool(use_graph),
                    l2=("L2 flushed between timed steps (512 MB written outside the per-step event pairs; state + tracks of a "
                        "rank: %.1f MB)" % (resident_bytes / 1e6)) if flush_l2 else
                       ("inputs larger than L2 (state + tracks of a rank: %.1f GB)" % (resident_bytes / 1e9))),
        roofline=roofline, cpu_baseline=cpu, gate_bounds=gate_bounds,
        e2e=dict(value=fu_e2e_all / e2e_sec, unit="feature-updates/s", h2d_bytes_per_step=h2d_bytes * world,
                 d2h_bytes_per_step=d2h_bytes * world, ms_per_step=ms_e2e / Ke, steps=Ke,
                 call="BatchedMSCKF.frame_step_detections (msckf_propagate / augment / match_descriptors / verify_matches / "
                      "append_* / update / select_poorest / drop_* / remove_cameras / compact_* _batched)",
                 inputs="per step and filter: 9 IMU samples + the frame's detections (256 x (keypoint, score, 32-d float32 "
                        "descriptor)) from pinned host memory, uploaded on a copy stream while the previous frame computes; "
                        "result: IMU state + window size back to pinned host memory",
                 filter_steps_per_s=B * world * Ke / e2e_sec),
        gpu_launches=int(launches), clocks=clocks, setup_s=setup_s, cpu_affinity_rank0=numa)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ncpu = os.cpu_count() or 1
    nw = min(ncpu, 64)
    for _ in range(0):
        pass
    t0 = time.perf_counter()
    steps = max(1, min(args.steps, 3))
    if args.config == 5:
        steps = 1
    run_cpu(min(nw, 4), 1, args.config if args.config != 5 else 4)    # warm-up: imports, page cache
    r = run_cpu(nw, steps, args.config)
    base = load_base_case(config=args.config)
    F, M, N = len(base["feat_len"]), base["obs_uv"].shape[1], int(base["n_cams"])
    sample = (f"{nw} filters x {steps} filter-step(s) each (9 propagations + {F}-feature update), one single-threaded "
              f"process per core; NumPy port of the reference update (the Python reference is not on the GPU box)")
    line = dict(impl="reference", metric="msckf_feature_updates_per_s", value=r["feature_updates_per_s"],
                unit="feature-updates/s", n_gpus=int(os.environ.get("WORLD_SIZE", "1")), steps=steps, warmup=1,
                ms_per_step=r["busy_s"] * 1e3 / steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f64", data="synthetic", filter_steps_per_s=r["filter_steps_per_s"],
                config=dict(workload=f"bounded sample of BASELINE.json {CONFIGS[args.config]['name']}: {F} features x {N}-pose window, "
                                     f"{N_IMU} IMU samples/frame, FP64; steady-state snapshot, step = 9 propagations + one update",
                            config=args.config, features=F, obs_per_feature=M, window=N),
                cpu_baseline=dict(value=r["feature_updates_per_s"], unit="feature-updates/s", cores=nw, kind="port",
                                  sample=sample),
                e2e=dict(value=r["feature_updates_per_s"], unit="feature-updates/s", h2d_bytes_per_step=0,
                         d2h_bytes_per_step=0),
                gpu_launches=0, wall_s=time.perf_counter() - t0)
    print(json.dumps(line))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--filters", type=int, default=4096, help="filters per GPU")
    ap.add_argument("--mode", default="gram", choices=["qr", "gram"],
                    help="compression of the stacked rows: gram = information form (default, fastest, same results to "
                         "1e-12), qr = explicit Householder TSQR as the reference's np.linalg.qr")
    ap.add_argument("--gate", default="exact", choices=["exact", "bounds"],
                    help="exact: every selected feature goes through the full Mahalanobis test (headline); bounds: same "
                         "decisions, features decided by the exact bounds skip the factorisation (reported beside it)")
    ap.add_argument("--window", type=int, default=30, choices=[30, 31],
                    help="camera poses in the window of the snapshot: 30 (BASELINE.json) or 31 (the transient max + 1 window "
                         "every other reference update sees, SURVEY.md section 8a)")
    ap.add_argument("--chunk", type=int, default=592,
                    help="filters per upload/compute/download chunk of the end-to-end (host-buffer) step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=4, choices=[2, 3, 4, 5],
                    help="BASELINE.json workload (1-based): 2 single filter, 3 = 64 filters, 4 = 4096 Monte-Carlo filters per GPU "
                         "(default, the configuration the metric is quoted on), 5 = stress 1024 x 1024 x 60 (128 filters per GPU)")
    ap.add_argument("--workload", default=None, choices=["stream", "snapshot"],
                    help="stream (default for configs 2-4): every filter on its own device-built stream, a step is a whole "
                         "camera frame; snapshot (config 5, and round 1's benchmark): one steady-state snapshot replicated with "
                         "per-filter noise, a step is 9 propagations + one update")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="stream workload: replay each camera frame from one captured CUDA graph (auto: up to 256 filters per GPU, "
                         "where the ~60 launches of a frame, not its kernels, set the frame time)")
    ap.add_argument("--profiler-range", action="store_true",
                    help="cudaProfilerStart/Stop around the timed steps of the stream workload (for ncu --profile-from-start off)")
    args = ap.parse_args()
    if "--filters" not in " ".join(sys.argv):
        args.filters = CONFIGS[args.config]["filters"]
    if args.workload is None:
        args.workload = "snapshot" if args.config == 5 else "stream"
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "stream":
        run_gpu_stream(args)
    else:
        run_gpu_snapshot(args)
