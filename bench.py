#!/usr/bin/env python
"""bench.py — differentiated frames/s of the CSFD/DCSFD KinectFusion frame loop on N B200s (one process per GPU).

  python bench.py --gpus 1 --steps K --warmup W            this repo's CUDA path (libxslam_b200.so, via the C-ABI)
  python bench.py --impl reference ...                      the CPU path timed on the box's host cores (oracle port)

A "step" is one ProcessFrame (surface measurement + 12 ICP iterations + TSDF integration + raycast + pyramid) on
the next frame of a synthetic ICL-NUIM-shaped sequence, producing the real outputs AND every requested derivative
direction.  Workload (BASELINE.json configs[2]/[3]): 640x480 depth, 512^3 TSDF @ 0.015 m, 55 DCSFD (bicomplex)
directions = 165 derivative components (the 21 axis pairs of the 6 pose DoF + 34 mixed pose-space directions; the
reference's intrinsics are real floats, so intrinsic directions are an extension that is not built yet).
With N > 1 the directions are sharded across ranks (strong scaling: total work fixed), every rank keeps a replica of
the real state, and the per-frame pose-derivative records are all-gathered over NCCL.
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--frames-per-step", type=int, default=10,
                    help="a step is one batch of this many consecutive depth frames (each one ProcessFrame); the timed region "
                         "is steps x frames-per-step frames so that it lasts ~1 s at the default flags")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--res", type=int, default=512)
    ap.add_argument("--dirs", type=int, default=55, help="bicomplex directions (mode dcsfd / hessian: pairs of the parameters) or first-order directions (csfd)")
    ap.add_argument("--comps", type=int, default=3, help="legacy selector: 1 = --mode csfd, 3 = the default DCSFD workload")
    ap.add_argument("--mode", default=None, choices=["hessian", "dcsfd", "csfd"],
                    help="how the DCSFD workload is carried: hessian (default) = Hessian-structured batch, n first-order + one "
                         "second-order plane per parameter pair (65 planes for the 55 pairs of 10 parameters); dcsfd = the same 55 "
                         "pairs as independent bicomplex directions (165 planes, round 1's layout); csfd = first-order directions")
    ap.add_argument("--pose-only", action="store_true",
                    help="hessian / dcsfd mode: the parameters beyond the 6 pose DoF are mixed pose-space directions instead of the "
                         "intrinsics fx, fy, cx, cy (round 1's workload)")
    ap.add_argument("--emulate-share", default=None, metavar="R/N",
                    help="hessian mode on one GPU: run rank R's share of an N-rank job (profiling aid; not a bench value)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sync-frames", action="store_true",
                    help="ProcessFrame waits for the end of each frame like the reference's (default: deferred mode, the "
                         "end-of-frame wait moves to the start of the next call)")
    ap.add_argument("--no-ref-cuda", action="store_true", help="skip timing the reference's own CUDA kernels (oracle/_ref)")
    return ap.parse_args()


def workload_cfg(xs, res):
    cfg = dict(xs.DEFAULT_CONFIG)
    cfg.update(tsdf_size_x=res, tsdf_size_y=res, tsdf_size_z=res, tsdf_voxel_size=7.68 / res)
    return cfg


def all_directions(xs, comps, dirs):
    """Seeds [dirs*comps, 16]: DCSFD -> 21 axis pairs (i<=j) of the pose DoF first, then deterministic mixed pairs."""
    if comps == 1:
        G = xs.se3_generators().reshape(6, 16)
        rng = np.random.default_rng(7)
        W = np.concatenate([np.eye(6), rng.standard_normal((max(dirs - 6, 0), 6)) / np.sqrt(6)])[:dirs]
        return (xs.H_ * W @ G).astype(np.float32)
    G = xs.se3_generators()
    rng = np.random.default_rng(7)
    pairs = [(np.eye(6)[i], np.eye(6)[j]) for i in range(6) for j in range(i, 6)]
    while len(pairs) < dirs:
        pairs.append((rng.standard_normal(6) / np.sqrt(6), rng.standard_normal(6) / np.sqrt(6)))
    out = np.zeros((dirs, 3, 16))
    for k, (u, w) in enumerate(pairs[:dirs]):
        Gu = np.tensordot(u, G, 1)
        Gw = np.tensordot(w, G, 1)
        out[k, 0] = (xs.H_ * Gu).reshape(16)
        out[k, 1] = (xs.H_ * Gw).reshape(16)
        out[k, 2] = (xs.H_ * xs.H_ * 0.5 * (Gu @ Gw + Gw @ Gu)).reshape(16)
    return out.reshape(-1, 16).astype(np.float32)


def hessian_params(dirs, intrinsics=True):
    """The parameters whose Hessian the DCSFD workload asks for: n with n (n + 1) / 2 = dirs (10 for 55 pairs).  BASELINE.json
    configs[3] names them: the 6 pose DoF (se3Exp coordinates) + the 4 intrinsics fx, fy, cx, cy.  With intrinsics=False the
    parameters beyond the sixth are deterministic mixed pose-space directions instead (round 1's workload).
    Returns (U [n, 6] pose-space directions, pairs, intrinsic seeds [n, 4] or None)."""
    n = int(round((np.sqrt(8 * dirs + 1) - 1) / 2))
    if n * (n + 1) // 2 != dirs:
        raise SystemExit("--mode hessian needs --dirs = n (n + 1) / 2 (21, 55, ...)")
    pairs = [(i, j) for i in range(n) for j in range(i, n)]
    if intrinsics and n > 6:
        U = np.concatenate([np.eye(6), np.zeros((n - 6, 6))])[:n]
        di = np.zeros((n, 4), np.float32)
        for p in range(6, min(n, 10)):
            di[p, p - 6] = H_STEP  # fx, fy, cx, cy
        return U, pairs, di
    rng = np.random.default_rng(7)
    U = np.concatenate([np.eye(6), rng.standard_normal((max(n - 6, 0), 6)) / np.sqrt(6)])[:n]
    return U, pairs, None


class ClockSampler(threading.Thread):
    """SM clock and clock-event (throttle) reasons sampled DURING the timed region (B200_PROFILING.md recipe).
    Uses NVML in-process (nvidia_ml_py) every 20 ms; falls back to polling `nvidia-smi` when NVML is unavailable."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.reasons, self.stop_flag, self.mx = index, [], set(), False, None
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
        for name, bit in bits.items():
            if r & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        r = [c.strip() for c in out.split(",")]
        if len(r) >= 7:
            self.sm.append(float(r[0]))
            self.mx = float(r[1])
            for i in range(4):
                if r[3 + i].lower().startswith("active"):
                    self.reasons.add(self.NAMES[i])

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            time.sleep(0.02 if self.nvml else 0.2)

    def summary(self):
        self.stop_flag = True
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx, "reasons": sorted(self.reasons),
                "samples": len(self.sm), "source": "nvml" if self.nvml else "nvidia-smi"}


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures of this same
# command (512^3, 55 directions, 1 GPU; the Hessian batch with its default parameters); None for any other configuration.
TRAFFIC_BY_MODE = {"dcsfd": {"icp_deriv": 1.243e9, "integrate": 6.40e8, "raycast_hit": 1.40e9},      # profiles/r01g_ncu_summary.md
                   "hessian": {"icp_deriv": 5.275e8, "integrate": 2.463e8, "raycast_hit": 5.204e8}}  # profiles/r02t_ncu_summary.md
TRAFFIC = {}


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ CPU baseline
# Everything in this section runs on the checker's side only (oracle/): it never imports the product package, so the
# reference arm's process maps no x-slam_b200/*.so.
REF_DEFAULTS = {  # Experiments/test_xkinect_fusion/configs/ICL_traj2.yaml:17-48 (the keys the frame loop reads)
    "biInterpolate_threshold": 0.0, "trunc_logistic_k": 0, "flag_use_gtPose": False, "max_integration_weight": 100, "thres_range": 3,
    "init_x": 3.2, "init_y": 3.2, "init_z": 3.2, "r_x": 0, "r_y": 0, "r_z": 0, "depth_width": 640, "depth_height": 480,
    "fx": 481.20, "fy": -480.00, "cx": 319.50, "cy": 239.50, "num_levels": 3, "distThres": 0.10, "angleThres": 15, "frame_step": 1}
H_STEP = 1e-7  # Internal.h:33


def ref_cfg(res):
    cfg = dict(REF_DEFAULTS)
    cfg.update(tsdf_size_x=res, tsdf_size_y=res, tsdf_size_z=res, tsdf_voxel_size=7.68 / res)
    return cfg


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm is meant to use every host core, so the value is set
    explicitly before the OpenMP runtime of liboracle.so starts, and again through the runtime's own API."""
    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)
    try:
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(cores)
    except OSError:
        pass
    return cores


def oracle_synth_depth(oracle, frame):
    """The synthetic stream restated on the checker's side (oracle/synth_oracle.cpp), pinned bit for bit against the
    product's generator by tests/test_bench_contract.py."""
    pose = np.zeros((16,), np.float32)
    oracle.lib.oracle_synth_pose(int(frame), pose.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    out = np.zeros((480, 640), np.uint16)
    c = REF_DEFAULTS
    oracle.lib.oracle_synth_depth(pose.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), ctypes.c_float(c["fx"]), ctypes.c_float(c["fy"]),
                                  ctypes.c_float(c["cx"]), ctypes.c_float(c["cy"]), 480, 640, out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)))
    return out


def cpu_port_frames(cfg, n_frames, budget_s=25.0):
    """Times the CPU oracle port of the frame loop (one first-order complex direction per run, exactly the
    reference's one-direction-per-run mode) with all host threads.  Returns the seconds of each frame-pass."""
    from oracle import pyref
    seed = np.zeros((4, 4), np.float32)
    seed[0, 3] = H_STEP  # world2camera(0,3) += i h: the commented seeding line KinectFusionReconstruction.cpp:22
    o = pyref.Oracle()
    k = pyref.OracleKinfu(cfg, seed, oracle=o)
    times = []
    t_all = time.perf_counter()
    for f in range(n_frames):
        d = oracle_synth_depth(o, f)
        t0 = time.perf_counter()
        ok = k.process_frame(d)
        times.append(time.perf_counter() - t0)
        if not ok or time.perf_counter() - t_all > budget_s:
            break
    return times


def test_csfd_baseline():
    """north_star's CPU baseline: the reference's host-side DeviceArray CSFD math as exercised by test_CSFD, on ONE host
    thread (the program is single-threaded): wall time of the unmodified Experiments/test_CSFD binary at -O0 (the reference's
    default build) and -O2, and consumed-result throughput of the unmodified DoubleComplex.cpp (oracle/_ref/libref_csfd.so):
    the test's own chain f1(t*t, sin t) and the element-wise ops * / sqrt sin."""
    from oracle import pyref
    out = {"kind": "reference", "cores": 1, "unit": "bicomplex chain evaluations/s",
           "what": "unmodified DeviceArray/src/DoubleComplex.cpp + Experiments/test_CSFD/main.cpp (oracle/_ref)"}
    for tag, path in (("O0", pyref.REF_TEST_CSFD), ("O2", pyref.REF_TEST_CSFD + "_O2")):
        if os.path.exists(path):
            t0 = time.perf_counter()
            r = subprocess.run([path], capture_output=True, text=True, timeout=120)
            out["test_CSFD_%s_wall_s" % tag] = time.perf_counter() - t0
            out["test_CSFD_%s_rc" % tag] = r.returncode
    if not os.path.exists(pyref.REF_CSFD_PATH):
        out["unavailable"] = "oracle/_ref/libref_csfd.so not built"
        return out
    ref = pyref.RefCsfd()
    rng = np.random.default_rng(0)
    n = 1 << 20
    t = rng.uniform(0.1, 1.5, n).astype(np.float32)
    rate, cs = ref.chain_bench(t, 1e-6, reps=4)
    out["value"] = rate
    out["sample"] = "4 x 2^20 evaluations of loss = f1(t*t, sin t) with seeded t (main.cpp:194-205), results summed (checksum %.6g)" % cs
    a = np.stack([rng.uniform(0.5, 2.0, n), 1e-6 * rng.standard_normal(n), 1e-6 * rng.standard_normal(n), 1e-12 * rng.standard_normal(n)], 1).astype(np.float32)
    b = np.stack([rng.uniform(0.5, 2.0, n), 1e-6 * rng.standard_normal(n), 1e-6 * rng.standard_normal(n), 1e-12 * rng.standard_normal(n)], 1).astype(np.float32)
    ops = {}
    for op in ("mul", "div", "sqrt", "sin"):
        t0 = time.perf_counter()
        ref.apply(op, a, b if op in ("mul", "div") else None)
        ops[op] = n / (time.perf_counter() - t0)
    out["ops_per_s"] = ops
    return out


def ref_cuda_frames(xs, cfg, n_frames=24):
    """The reference's own CUDA kernels (oracle/_ref/libxslam_ref.so: unmodified XKinectFusion/src/*.cu recompiled for
    sm_100a, driven by the restated orchestrator) on the same GPU: one first-order complex direction per pass, exactly
    how the reference would produce k directions (k passes).  Returns seconds per pass-frame (frames 1.. only)."""
    from oracle import pyref
    if not os.path.exists(pyref.REF_CUDA_PATH):
        return None
    ref = pyref.RefCuda()
    r = ref.kinfu(cfg, xs.pose_seeds_csfd()[0].reshape(4, 4))  # imaginary part of world2camera (KFR.cpp:22)
    times = []
    for f in range(n_frames):
        d = xs.synth_depth(f)
        t0 = time.perf_counter()
        ok = r.process_frame(d)  # uploads the frame, runs ProcessFrame with the reference's own syncs
        times.append(time.perf_counter() - t0)
        if not ok:
            break
    del r
    return times


def distinct_planes(args):
    """Derivative planes a differentiated frame of the workload consists of, counted once: the DCSFD Hessian of n parameters
    has n first-order and n(n+1)/2 second-order components (65 at n = 10, i.e. 55 bicomplex directions)."""
    if args.comps == 1 or args.mode == "csfd":
        return args.dirs
    n = int(round((np.sqrt(8 * args.dirs + 1) - 1) / 2))
    return n + args.dirs if n * (n + 1) // 2 == args.dirs else 3 * args.dirs


def run_reference(args, rank):
    """--impl reference: the CPU path (oracle port of the reference's frame loop; the reference itself has no CPU
    implementation of this path and its orchestrator cannot be built offline) on all host cores.  A step is ONE pass of the
    frame loop over one frame with ONE complex perturbation direction - the reference's own mode of operation (one
    imaginary part per run) and a bounded sample of the workload.  steps / ms_per_step are what was timed; `value` scales
    the measured pass time by the number of passes a differentiated frame needs (every distinct derivative plane is at
    least one pass), stated in `passes_per_differentiated_frame`."""
    if rank != 0:
        return
    cores = use_all_host_threads()
    cfg = ref_cfg(args.res)
    total = args.warmup + args.steps
    times = cpu_port_frames(cfg, total, budget_s=240.0)
    timed = times[args.warmup:] if len(times) > args.warmup else times[-1:]
    per_pass = float(np.mean(timed))
    passes = distinct_planes(args)
    fps = 1.0 / (per_pass * passes)
    sample = ("%d frame-passes of the 640x480 / %d^3 sequence timed, ONE first-order complex direction per pass (the reference's "
              "one-direction-per-run mode, %d host threads); value = 1 / (%d passes per differentiated frame x measured pass time)"
              % (len(timed), args.res, cores, passes))
    line = {"impl": "reference", "metric": "differentiated_frames_per_s", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": len(timed), "warmup": min(args.warmup, len(times) - len(timed)), "ms_per_step": per_pass * 1e3,
            "step_is": "one frame-pass with one complex direction (bounded sample)", "passes_per_differentiated_frame": passes,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "icl_synth_640x480_tsdf%d_dcsfd%d" % (args.res, args.dirs)},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ CUDA path
BATCH_NOTE = {"hessian": "Hessian-structured batch (comps = 2): n first-order planes + one second-order plane per parameter pair",
              "dcsfd": "DCSFD list (comps = 3): every pair an independent bicomplex direction (eps1, eps2, eps1eps2)",
              "csfd": "CSFD list (comps = 1): independent first-order directions"}
ICP_KERNEL = {"hessian": "icp_deriv_tile_kernel", "dcsfd": "icp_deriv_kernel<3>", "csfd": "icp_deriv_kernel<1>"}


class DeviceRecord:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


def run_ours(args, xs, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    cfg = workload_cfg(xs, args.res)
    mode = args.mode
    k = xs.KinectFusionReconstruction()
    max_dirs = (args.dirs + world - 1) // world
    if mode == "hessian":
        # every rank carries the n first-order components (cheap, and every pair needs two of them) and its share of the pairs
        # blocked shards (parallel.plan_hessian_shards): a rank carries its block of the pairs and the first-order components of
        # the parameters those pairs touch - the real state is replicated
        from xslam_b200 import parallel as par
        U, pairs, dintr = hessian_params(args.dirs, intrinsics=not args.pose_only)
        n_params = U.shape[0]
        share_rank, share_world = rank, world
        if args.emulate_share:  # one GPU running rank r's share of an N-rank job (profiling aid)
            share_rank, share_world = (int(t) for t in args.emulate_share.split("/"))
        plan = par.plan_hessian_shards(n_params, pairs, share_world)
        mine = plan[share_rank]
        my_seeds, _ = xs.hessian_seeds(U[mine["params"]], mine["local_pairs"])
        my_dintr = None if dintr is None else np.ascontiguousarray(dintr[mine["params"]])
        if my_dintr is not None and not my_dintr.any():
            my_dintr = None
        k.SetYamlParameters(cfg, comps=2, seeds=my_seeds, pairs=mine["local_pairs"], n_params=len(mine["params"]), intrinsic_seeds=my_dintr)
        ncomp_local, ncomp_max = len(mine["params"]) + len(mine["pair_ids"]), par.planned_record_floats(plan) // 16 - 1
        planes_total = n_params + len(pairs)
    else:
        comps = 1 if mode == "csfd" else 3
        if mode == "dcsfd":  # the pairs of the same parameters as independent bicomplex directions (eps1, eps2, eps1eps2)
            U, pairs, _ = hessian_params(args.dirs, intrinsics=False)  # a DCSFD list cannot carry intrinsic parameters
            G = np.tensordot(U, xs.se3_generators(), 1)
            seeds = np.zeros((args.dirs, 3, 16))
            for d, (i, j) in enumerate(pairs):
                seeds[d, 0], seeds[d, 1] = (xs.H_ * G[i]).reshape(16), (xs.H_ * G[j]).reshape(16)
                seeds[d, 2] = (xs.H_ * xs.H_ * 0.5 * (G[i] @ G[j] + G[j] @ G[i])).reshape(16)
            seeds = seeds.astype(np.float32)
        else:
            seeds = all_directions(xs, 1, args.dirs).reshape(args.dirs, 1, 16)
        mine = list(range(rank, args.dirs, world))
        k.SetYamlParameters(cfg, comps=comps, seeds=np.ascontiguousarray(seeds[mine].reshape(-1, 16)))
        ncomp_local, ncomp_max = len(mine) * comps, max_dirs * comps
        planes_total = args.dirs * comps
    lib = xs.load()
    rec_len = (1 + ncomp_max) * 16
    comm = None
    if world > 1:
        # the multi-GPU layer is the library's own (csrc/comm.cpp): every ProcessFrame queues one NCCL all-gather of the ranks'
        # pose records on an internal stream behind the frame's record upload; torch.distributed only hands out the NCCL id
        from xslam_b200 import parallel
        comm = parallel.Comm.from_torch_distributed(dist, device="cuda")
        k.set_comm(comm, rec_len)

    # A step is one batch of FPS consecutive frames (each one ProcessFrame).  The synthetic trajectory is a closed loop of
    # period 300 frames, so frame f of the stream is frame f % 300 of the generator: at most 300 distinct frames are rendered
    # on the host, resident copies live in HBM (timed region 1) and in pinned host memory (timed region 2).
    W, K, FPS = args.warmup, args.steps, max(1, args.frames_per_step)
    n_stream = (W + 2 * K) * FPS
    PERIOD = 300
    frames = [xs.synth_depth(f) for f in range(min(n_stream, PERIOD))]
    dev_frames = [torch.from_numpy(f.astype(np.int16)).cuda() for f in frames]
    pinned = [torch.from_numpy(f.astype(np.int16)).pin_memory() for f in frames]
    frame_no = [0]  # index of the next frame of the stream

    lib_stream = torch.cuda.ExternalStream(k.stream_ptr())
    deferred = not args.sync_frames
    if deferred:
        k.set_deferred(True)

    def step(depth):
        ok = k.ProcessFrame(depth)
        if not ok:
            raise RuntimeError("frame alignment failed: " + lib.xs_last_error().decode())

    def sync():
        k.sync()  # collects a deferred frame and waits for its all-gather
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def next_frame(pool):
        f = pool[frame_no[0] % len(pool)]
        frame_no[0] += 1
        return f

    for i in range(W * FPS):
        step(next_frame(dev_frames))
    # ---------------- timed region 1: inputs resident in HBM
    sampler = ClockSampler(local_rank)
    sampler.start()
    stage_ms = {n: 0.0 for n in ("surface", "icp", "integrate", "raycast", "total")}
    abytes = {n: 0.0 for n in ("surface", "icp", "integrate", "raycast")}
    kern_ms, upd = 0.0, 0
    hit_ms, hits, normals = 0.0, 0, 0
    rstats = (ctypes.c_ulonglong * 2)()
    icp_ms, icp_n = 0.0, 0
    t_ms = (ctypes.c_float * 16)()
    t_px = (ctypes.c_int * 16)()
    vol = lib.xs_kinfu_volume(k.h)
    sync()
    l0 = lib.xs_launch_count()
    # CUDA events on the stream the library launches on (torch.cuda.Event sees only the stream it is recorded on)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(lib_stream)
    t0 = time.perf_counter()
    def collect():  # stage times / statistics of the last collected frame
        nonlocal kern_ms, upd, hit_ms, hits, normals
        hit_ms += lib.xs_volume_last_raycast_hit_ms(vol)
        lib.xs_volume_raycast_stats(vol, rstats)
        hits += rstats[0]
        normals += rstats[1]
        tm, _ = k.times()
        for n in stage_ms:
            stage_ms[n] += tm[n]
        ab = k.algorithmic_bytes()
        for n in abytes:
            abytes[n] += ab[n]
        kern_ms += lib.xs_volume_last_integrate_ms(vol)
        upd += k.stats()[0]

    NF = K * FPS  # frames in each timed region
    for i in range(NF):
        step(next_frame(dev_frames))
        # deferred mode: frame i is still integrating / raycasting here; the getters describe frame i - 1 (collected at the
        # start of this step), so the per-stage sums lag by one frame and the last frame is collected after the final sync
        if not deferred or i > 0:
            collect()
        for j in range(lib.xs_icp_deriv_times(t_ms, t_px, 16)):  # level-0 launches of the dominant kernel (ICP is complete)
            if t_px[j] == 640 * 480:
                icp_ms += t_ms[j]
                icp_n += 1
    ev1.record(lib_stream)
    sync()
    t_wall = time.perf_counter() - t0
    if deferred:
        collect()
    # the frame loop synchronises with the host at least once per frame (the ICP result), so the device-event bracket and
    # the wall clock agree;
    # the reported value uses the device events (max over ranks below)
    t_dev = ev0.elapsed_time(ev1) * 1e-3
    launches = lib.xs_launch_count() - l0
    # ---------------- timed region 2: end to end through the public call with HOST buffers
    sync()
    t0 = time.perf_counter()
    for i in range(NF):
        step(next_frame(pinned).numpy().view(np.uint16))
        # the frame's result, read on the host every frame.  One rank: the pose record (value + derivative components).  N ranks:
        # all ranks' records; the consumer runs one frame behind - after submitting frame i it reads frame i - 1's gathered
        # records, whose all-gather ran beside frame i's kernels - and reads the last frame's after the loop
        if comm is None:
            w2c = k.world2camera
        elif i > 0:
            w2c = k.gathered_records(lag=1)
    if comm is not None:
        w2c = k.gathered_records()
    sync()
    t_e2e = time.perf_counter() - t0
    clocks = sampler.summary()
    if world > 1:
        tt = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_e2e = float(tt[0]), float(tt[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    if world == 1 and args.res == 512 and args.dirs == 55 and not (mode == "hessian" and args.pose_only):
        TRAFFIC.update(TRAFFIC_BY_MODE.get(mode, {}))
    peak, peak_src = measured_hbm_peak()
    int_bytes = abytes["integrate"] / NF
    int_ms = kern_ms / NF
    achieved = int_bytes / (int_ms * 1e-3) / 1e9 if int_ms > 0 else 0.0
    # dominant kernel of the step: icp_deriv_kernel at pyramid level 0 (5 launches per frame).  Algorithmic bytes per
    # launch: per pixel the real current + previous maps (48 B) and 24 B per derivative
    # component of the previous maps, plus the 27 sums per component written out.
    D = ncomp_local
    icp_bytes = 640 * 480 * (48 + 24 * D) + 27 * 8 * (1 + D)
    icp_kernel_ms = icp_ms / icp_n if icp_n else 0.0
    icp_achieved = icp_bytes / (icp_kernel_ms * 1e-3) / 1e9 if icp_kernel_ms > 0 else 0.0
    # raycast hit kernel: its unavoidable traffic is the output maps (24 B per pixel and component incl. the real one) plus
    # the hit-time image; the trilinear gathers (SURVEY 8d counts 64 x 4 B per hit pixel and component, before any cache
    # reuse between neighbouring pixels) are data dependent and reported separately
    hit_kernel_ms = hit_ms / NF
    hit_bytes = 640 * 480 * (24 * (1 + D) + 4)
    hit_achieved = hit_bytes / (hit_kernel_ms * 1e-3) / 1e9 if hit_kernel_ms > 0 else 0.0
    kernels = {"icp_deriv (pyramid level 0, 5 launches per frame)": 5 * icp_kernel_ms, "raycast_hit": hit_kernel_ms, "integrate": int_ms}
    dominant = max(kernels, key=kernels.get)
    # per step: depth frame + pose derivative components for ICP (initial pose), integration (v2c) and raycast
    # (c2v, v2w) in; final ICP pose with all derivative components, status and integration statistics out
    h2d = 640 * 480 * 2 + (1 + ncomp_local) * 48 + ncomp_local * 48 * 3 + (1 + ncomp_local) * 64
    d2h = (1 + ncomp_local) * 48 + 8 + 32
    line = {
        "metric": "differentiated_frames_per_s", "value": NF / t_dev, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": t_dev / K * 1e3, "frames_per_step": FPS, "ms_per_frame": t_dev / NF * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "icl_synth_640x480_tsdf%d_dcsfd%d" % (args.res, args.dirs), "step": "one batch of %d consecutive depth frames, each one ProcessFrame" % FPS,
                   "depth": "640x480 uint16 mm",
                   "tsdf": "%d^3 @ %.4f m" % (args.res, 7.68 / args.res), "directions": args.dirs, "batch": BATCH_NOTE[mode],
                   "parameters": ("6 pose DoF + fx, fy, cx, cy" if (mode == "hessian" and not args.pose_only and args.dirs == 55)
                                  else "pose-space directions") if mode != "csfd" else "pose-space directions",
                   "derivative_planes": planes_total, "derivative_planes_rank0": ncomp_local, "directions_per_rank": max_dirs,
                   "sharding": "blocks of the second-order pairs over ranks, each with the first-order components its pairs touch; real state replicated" if mode == "hessian"
                               else "directions over ranks, real state replicated",
                   "frame_sync": "deferred (end-of-frame wait at the start of the next ProcessFrame; ICP result read on the host every frame)" if deferred else "every frame",
                   "l2": "per-step working set (volume %.1f GB/rank) >> 126 MB L2, no flush needed" % (lib.xs_volume_bytes(vol) / 1e9)},
        "e2e": {"value": NF / t_e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d * FPS, "d2h_bytes_per_step": d2h * FPS},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "%s (pyramid level 0; 5 launches = %.0f%% of the frame)" % (ICP_KERNEL[mode], 100 * 5 * icp_kernel_ms / (t_dev / NF * 1e3)),
                     "achieved": icp_achieved, "peak": peak, "unit": "GB/s", "frac": icp_achieved / peak, "traffic": TRAFFIC.get("icp_deriv"),
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": icp_bytes, "kernel_ms": icp_kernel_ms, "launches_timed": icp_n,
                     "bytes_model": "P0 x (48 + 24 D) + 27 x 8 x (1 + D), D = derivative planes, each counted once (SURVEY 8d)",
                     "co_bound_note": "instruction issue at 11 warps per SM and the L1 capacity left beside 158 KB of staging bound this kernel before HBM does (DESIGN.md 5.2)"},
        "roofline_integrate": {"bound": "hbm", "kernel": "integrate_kernel<%s>" % mode, "achieved": achieved, "peak": peak, "unit": "GB/s",
                               "frac": achieved / peak, "traffic": TRAFFIC.get("integrate"), "algorithmic_bytes_per_launch": int_bytes,
                               "kernel_ms": int_ms, "updated_voxels_per_launch": upd / NF,
                               "survey_8d_bytes_per_launch": (upd / NF) * 2 * (8 + 4 * D) + 2 * 640 * 480},
        "roofline_raycast": {"bound": "hbm", "kernel": "raycast_hit_kernel<%s>" % mode, "achieved": hit_achieved, "peak": peak, "unit": "GB/s",
                             "frac": hit_achieved / peak, "traffic": TRAFFIC.get("raycast_hit"), "algorithmic_bytes_per_launch": hit_bytes,
                             "kernel_ms": hit_kernel_ms, "hit_pixels_per_launch": hits / NF, "normal_pixels_per_launch": normals / NF,
                             "bytes_model": "output maps 24 B x (1 + D) per pixel + hit-time image; gathers excluded (data dependent)",
                             "survey_8d_gather_bytes_per_launch": (hits / NF) * 64 * 4 * (1 + D)},
        "kernel_ms_per_frame": kernels, "dominant_kernel": dominant,
        "wall_ms_per_frame": t_wall / NF * 1e3,
        "stages_ms_per_frame": {n: v / NF for n, v in stage_ms.items()},
        "stages_note": ("icp / integrate / raycast: CUDA-event brackets on the pipeline's stream; surface: the next frame's head runs "
                        "on the second stream beside the previous frame's raycast (deferred mode), total = sum of the four") if deferred
                       else "CUDA-event brackets on the pipeline's stream",
        "stages_algorithmic_GBps": {n: abytes[n] / (stage_ms[n] * 1e-3) / 1e9 if stage_ms[n] > 0 else 0.0 for n in abytes},
        "frame_algorithmic_bytes": sum(abytes.values()) / NF,
    }
    if world == 1 and not args.no_cpu_baseline:
        cores = use_all_host_threads()
        times = cpu_port_frames(ref_cfg(args.res), 40, budget_s=20.0)
        per_dir = float(np.mean(times[1:])) if len(times) > 1 else float(times[0])
        ncd = distinct_planes(args)
        line["cpu_baseline"] = {"value": 1.0 / (per_dir * ncd), "unit": "frames/s", "cores": cores, "kind": "port",
                                "sample": "%d frame-passes timed, 640x480 / %d^3, ONE first-order complex direction per pass (reference's "
                                          "one-direction-per-run mode, all host threads); value = 1 / (%d passes per differentiated frame "
                                          "x measured pass time)" % (max(len(times) - 1, 1), args.res, ncd),
                                "seconds_per_direction_frame": per_dir, "passes_per_differentiated_frame": ncd}
        # north_star's own CPU baseline: the host DeviceArray CSFD math of test_CSFD (unmodified reference sources), beside the
        # same chain on the packed-SoA device arrays (xs_dc_chain)
        tc = test_csfd_baseline()
        try:
            from xslam_b200 import ops
            n = 1 << 24
            tt = torch.rand(n, device="cuda") * 1.4 + 0.1
            for _ in range(3):
                ops.dc_chain(tt, 1e-6)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                ops.dc_chain(tt, 1e-6)
            e1.record()
            torch.cuda.synchronize()
            tc["gpu_value"] = 10 * n / (e0.elapsed_time(e1) * 1e-3)
            tc["gpu_sample"] = "10 x 2^24 evaluations of the same chain on packed-SoA device arrays (xs_dc_chain), inputs resident in HBM"
        except Exception as e:
            tc["gpu_value"] = None
            tc["gpu_error"] = str(e)[:200]
        line["cpu_baseline_test_csfd"] = tc
    if world == 1 and not args.no_ref_cuda and not args.no_cpu_baseline:
        del k
        torch.cuda.empty_cache()
        try:
            rt = ref_cuda_frames(xs, cfg)
        except Exception as e:  # the checker is optional on the box
            rt = None
            line["ref_cuda_baseline"] = {"unavailable": str(e)[:200]}
        if rt:
            per = float(np.mean(rt[1:])) if len(rt) > 1 else float(rt[0])
            ncd = distinct_planes(args)
            line["ref_cuda_baseline"] = {
                "value": 1.0 / (per * ncd), "unit": "frames/s", "kind": "reference CUDA kernels recompiled for sm_100a (oracle/_ref/libxslam_ref.so), same B200",
                "seconds_per_direction_frame": per, "passes_per_differentiated_frame": ncd,
                "sample": "%d frames timed, 640x480 / %d^3, one first-order complex direction per pass; value = 1 / (%d passes x measured pass time)" % (max(len(rt) - 1, 1), args.res, ncd)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.mode is None:
        args.mode = "csfd" if args.comps == 1 else "hessian"
        if args.mode == "hessian":  # a direction count that is not a full pair set runs as a plain DCSFD list
            n = int(round((np.sqrt(8 * args.dirs + 1) - 1) / 2))
            if n * (n + 1) // 2 != args.dirs:
                args.mode = "dcsfd"
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":  # never imports the product
        run_reference(args, rank)
        return
    import xslam_b200 as xs
    run_ours(args, xs, rank, world, local_rank)


if __name__ == "__main__":
    main()
