"""Host-side mirror of class KinectFusionReconstruction (XKinectFusion/include/KinectFusionReconstruction.h:19-220)
over the C-ABI pipeline object (xs_kinfu).  Method names, the YAML keys and the 1/0 return convention are the
reference's; the perturbation seeds generalise the commented seeding line KinectFusionReconstruction.cpp:22
to a batch of k directions (CSFD, comps=1) or k bicomplex directions (DCSFD, comps=3).
"""
import ctypes as C
import math
import os

import numpy as np

from . import _capi
from ._capi import Config, check

H_ = 1e-7  # Internal.h:33


def load_yaml(path):
    """Flat `key: value` YAML reader (yaml-cpp is not available; the reference config is flat:
    Experiments/test_xkinect_fusion/configs/ICL_traj2.yaml)."""
    cfg = {}
    with open(path) as f:
        for line in f:
            line = line.split("#", 1)[0].strip()
            if not line or ":" not in line:
                continue
            k, v = line.split(":", 1)
            v = v.strip().strip('"').strip("'")
            if v.lower() in ("true", "false"):
                cfg[k.strip()] = v.lower() == "true"
                continue
            try:
                cfg[k.strip()] = int(v)
            except ValueError:
                try:
                    cfg[k.strip()] = float(v)
                except ValueError:
                    cfg[k.strip()] = v
    return cfg


DEFAULT_CONFIG = {  # Experiments/test_xkinect_fusion/configs/ICL_traj2.yaml:17-48
    "biInterpolate_threshold": 0.0, "trunc_logistic_k": 0, "flag_use_gtPose": False,
    "tsdf_size_x": 256, "tsdf_size_y": 256, "tsdf_size_z": 256, "tsdf_voxel_size": 0.03,
    "max_integration_weight": 100, "thres_range": 3, "init_x": 3.2, "init_y": 3.2, "init_z": 3.2,
    "r_x": 0, "r_y": 0, "r_z": 0, "depth_width": 640, "depth_height": 480,
    "fx": 481.20, "fy": -480.00, "cx": 319.50, "cy": 239.50, "num_levels": 3, "distThres": 0.10, "angleThres": 15,
    "frame_step": 1,
}


def make_config(cfg):
    c = Config()
    c.res[:] = [int(cfg["tsdf_size_x"]), int(cfg["tsdf_size_y"]), int(cfg["tsdf_size_z"])]
    c.voxel_size = float(cfg["tsdf_voxel_size"])
    c.max_weight = int(cfg["max_integration_weight"])
    c.thres_range = float(cfg["thres_range"])
    c.init_xyz[:] = [float(cfg["init_x"]), float(cfg["init_y"]), float(cfg["init_z"])]
    c.r_deg[:] = [float(cfg["r_x"]), float(cfg["r_y"]), float(cfg["r_z"])]
    c.width, c.height = int(cfg["depth_width"]), int(cfg["depth_height"])
    c.fx, c.fy, c.cx, c.cy = float(cfg["fx"]), float(cfg["fy"]), float(cfg["cx"]), float(cfg["cy"])
    c.num_levels = int(cfg["num_levels"])
    c.dist_thres = float(cfg["distThres"])
    c.angle_thres_deg = float(cfg["angleThres"])
    c.bi_threshold = float(cfg["biInterpolate_threshold"])
    c.trunc_k = float(cfg["trunc_logistic_k"])
    c.frame_step = int(cfg.get("frame_step", 1))  # KinectFusionReconstruction.cpp:72
    return c


def se3_generators():
    """d/d(xi_i) of se3Exp(xi) at xi = 0 (KinectFusionReconstruction.h:176-219, xi = [v; omega]) as 4x4 matrices."""
    G = np.zeros((6, 4, 4), np.float64)
    for i in range(3):
        G[i, i, 3] = 1.0
    G[3, 1, 2], G[3, 2, 1] = -1.0, 1.0
    G[4, 0, 2], G[4, 2, 0] = 1.0, -1.0
    G[5, 0, 1], G[5, 1, 0] = -1.0, 1.0
    return G


def pose_seeds_csfd(h=H_):
    """6 CSFD directions: world2camera = se3Exp(h e_i) -> imaginary part h * G_i.  [6, 16] float32."""
    return (h * se3_generators()).reshape(6, 16).astype(np.float32)


def pose_seeds_dcsfd(pairs=None, h=H_, n_params=6):
    """DCSFD directions for parameter pairs (i, j): eps1 along e_i, eps2 along e_j; the second-order seed
    is h^2 * (G_i G_j + G_j G_i)/2 (second derivative of se3Exp along the two axes).  [k*3, 16] float32."""
    G = se3_generators()
    if pairs is None:
        pairs = [(i, j) for i in range(n_params) for j in range(i, n_params)]
    out = np.zeros((len(pairs), 3, 16), np.float64)
    for k, (i, j) in enumerate(pairs):
        out[k, 0] = (h * G[i]).reshape(16)
        out[k, 1] = (h * G[j]).reshape(16)
        out[k, 2] = (h * h * 0.5 * (G[i] @ G[j] + G[j] @ G[i])).reshape(16)
    return out.reshape(-1, 16).astype(np.float32), pairs


def all_pairs(n):
    """The n (n + 1) / 2 index pairs (i <= j) of a Hessian batch, in the order the library lists them by default."""
    return [(i, j) for i in range(n) for j in range(i, n)]


def hessian_seeds(params=None, pairs=None, h=H_):
    """Seeds of a Hessian batch (comps = 2) for pose-space parameters: params [n, 6] are se3Exp coordinate directions
    (default: the 6 axes), parameter p perturbs world2camera along G_p = sum_a params[p, a] G_a.  Returns
    (seeds [(n + m), 16] float32, pairs): rows 0..n-1 = h G_p (first order), then h^2 (G_i G_j + G_j G_i) / 2 per pair - the
    (eps1, eps2, eps1eps2) seeds of pose_seeds_dcsfd with every first-order seed stored once."""
    G = se3_generators()
    U = np.eye(6) if params is None else np.asarray(params, np.float64).reshape(-1, 6)
    Gp = np.tensordot(U, G, 1)
    n = Gp.shape[0]
    if pairs is None:
        pairs = all_pairs(n)
    out = np.zeros((n + len(pairs), 16), np.float64)
    for p in range(n):
        out[p] = (h * Gp[p]).reshape(16)
    for k, (i, j) in enumerate(pairs):
        out[n + k] = (h * h * 0.5 * (Gp[i] @ Gp[j] + Gp[j] @ Gp[i])).reshape(16)
    return out.astype(np.float32), list(pairs)


def se3_exp(xi, comps=1, dirs=None, pairs=None):
    """se3Exp (KinectFusionReconstruction.h:176-219) on batched jets.  xi: [(1 + ncomp), 6] = (v, omega) with component 0 real
    and the rest h-scaled derivative components (comps / dirs / pairs as for SetYamlParameters).  Returns [(1 + ncomp), 4, 4]."""
    x = np.ascontiguousarray(xi, np.float32).reshape(-1, 6)
    ncomp = x.shape[0] - 1
    if comps == 2:
        if dirs is None:
            dirs = ncomp - len(pairs) if pairs is not None else int(round((np.sqrt(9 + 8 * ncomp) - 3) / 2))
        plist = all_pairs(dirs) if pairs is None else list(pairs)
        pa = np.ascontiguousarray(np.asarray(plist, np.int32).reshape(-1, 2))
        pp, npairs = pa.ctypes.data_as(C.POINTER(C.c_int)), len(plist)
    else:
        dirs = ncomp // comps if dirs is None else dirs
        pp, npairs = None, 0
    out = np.zeros((1 + ncomp, 16), np.float32)
    check(_capi.load().xs_se3_exp(x.ctypes.data_as(C.POINTER(C.c_float)), comps, dirs, npairs, pp, out.ctypes.data_as(C.POINTER(C.c_float))),
          "se3_exp")
    return out.reshape(-1, 4, 4)


class _DeviceView:
    """Zero-copy view of library-owned device memory through the CUDA array interface."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


class KinectFusionReconstruction:
    """Drop-in for the reference class on the hot path.  depth frames are uint16 millimetres [H, W]."""

    SOLVE_EIGEN_LLT, SOLVE_ANALYTIC = 0, 1

    def __init__(self):
        self.lib = _capi.load()
        self.h = None
        self.depth_width = 0
        self.depth_height = 0
        self.config = None

    def SetYamlParameters(self, config, comps=1, seeds=None, solve_mode=None, pairs=None, n_params=None, intrinsic_seeds=None,
                          keep_current_map_derivatives=False):
        """KinectFusionReconstruction.cpp:9-73.  config: dict of YAML keys (or a path).
        seeds: [ncomp, 16] h-scaled derivative components of the initial world2camera.  comps = 1 / 3: lists of first-order /
        bicomplex directions (ncomp = dirs * comps).  comps = 2: Hessian batch over n_params parameters - n_params first-order
        rows followed by one second-order row per pair of `pairs` (default: all pairs; see hessian_seeds)."""
        if isinstance(config, str):
            config = load_yaml(config)
        cfg = dict(DEFAULT_CONFIG)
        cfg.update(config)
        self.config = cfg
        if cfg["num_levels"] > 3:
            print("sorry, the max supported multi-level = 3")
        self.depth_width, self.depth_height = int(cfg["depth_width"]), int(cfg["depth_height"])
        self.frame_step = int(cfg.get("frame_step", 1))
        seeds = None if seeds is None else np.ascontiguousarray(seeds, np.float32).reshape(-1, 16)
        ncomp = 0 if seeds is None else seeds.shape[0]
        self.pairs = None
        if comps == 2:
            if n_params is None:  # all pairs: ncomp = n + n (n + 1) / 2
                if pairs is not None:
                    n_params = ncomp - len(pairs)
                else:
                    n_params = int(round((np.sqrt(9 + 8 * ncomp) - 3) / 2))
            self.pairs = all_pairs(n_params) if pairs is None else [tuple(int(v) for v in p) for p in pairs]
            if n_params + len(self.pairs) != ncomp:
                raise ValueError("a Hessian batch needs n_params + len(pairs) seed rows")
            self.comps, self.dirs, self.ncomp = 2, n_params, ncomp
        else:
            if ncomp % comps:
                raise ValueError("seeds must hold dirs*comps rows")
            self.comps, self.dirs, self.ncomp = comps, ncomp // comps, ncomp
        if solve_mode is None:
            solve_mode = self.SOLVE_EIGEN_LLT if comps == 1 else self.SOLVE_ANALYTIC
        c = make_config(cfg)
        sp = seeds.ctypes.data_as(C.POINTER(C.c_float)) if ncomp else None
        if self.h:
            self.lib.xs_kinfu_destroy(self.h)
        if comps == 2:
            pa = np.ascontiguousarray(np.asarray(self.pairs, np.int32).reshape(-1, 2))
            self.h = self.lib.xs_kinfu_create_hessian(C.byref(c), self.dirs, len(self.pairs), pa.ctypes.data_as(C.POINTER(C.c_int)), sp,
                                                      solve_mode)
        else:
            self.h = self.lib.xs_kinfu_create(C.byref(c), comps, self.dirs, sp, solve_mode)
        if not self.h:
            raise _capi.XsError("SetYamlParameters: " + self.lib.xs_last_error().decode())
        if intrinsic_seeds is not None:
            # [n_params, 4]: h d(fx, fy, cx, cy) / d theta_p - parameters of a Hessian batch that move the intrinsics
            d = np.ascontiguousarray(intrinsic_seeds, np.float32).reshape(-1, 4)
            if comps != 2 or d.shape[0] != self.dirs:
                raise ValueError("intrinsic_seeds: [n_params, 4] rows for a Hessian batch (comps = 2)")
            check(self.lib.xs_kinfu_set_intrinsic_seeds(self.h, d.ctypes.data_as(C.POINTER(C.c_float))), "intrinsic_seeds")
            if keep_current_map_derivatives:
                check(self.lib.xs_kinfu_keep_current_map_derivatives(self.h, 1), "keep_current_map_derivatives")
        self.use_gtPose = bool(cfg.get("flag_use_gtPose", False))  # KinectFusionReconstruction.cpp:69
        self._gt_poses = []                                        # :70 gt_poses.resize(0)
        if self.use_gtPose:
            self.lib.xs_kinfu_set_gt_poses(self.h, None, 0, 1)

    @property
    def gt_poses(self):
        """KinectFusionReconstruction.h:36: camera-to-world ground-truth poses, used when flag_use_gtPose is set."""
        return self._gt_poses

    @gt_poses.setter
    def gt_poses(self, poses):
        self._gt_poses = [np.asarray(p, np.float32).reshape(4, 4) for p in poses]
        flat = np.ascontiguousarray(np.stack(self._gt_poses), np.float32) if self._gt_poses else np.zeros((0, 16), np.float32)
        _capi.check(self.lib.xs_kinfu_set_gt_poses(self.h, flat.ctypes.data_as(C.POINTER(C.c_float)), len(self._gt_poses),
                                                  int(self.use_gtPose)), "gt_poses")

    def ReleaseBuffers(self):
        if self.h:
            self.lib.xs_kinfu_destroy(self.h)
            self.h = None

    __del__ = ReleaseBuffers

    # -- the frame loop -------------------------------------------------------------------------
    def ProcessFrame(self, depth):
        """KinectFusionReconstruction.cpp:147-159.  depth: numpy uint16 [H, W] (host) or a torch CUDA int16/uint16
        tensor (device resident).  Returns 1, or 0 when frame alignment failed."""
        if hasattr(depth, "data_ptr"):
            return self.lib.xs_kinfu_process_frame(self.h, C.c_void_p(depth.data_ptr()), 1)
        d = np.ascontiguousarray(depth, np.uint16)
        if d.shape != (self.depth_height, self.depth_width):
            raise ValueError("depth frame must be [%d, %d]" % (self.depth_height, self.depth_width))
        return self.lib.xs_kinfu_process_frame(self.h, d.ctypes.data_as(C.c_void_p), 0)

    def set_deferred(self, on=True):
        """Throughput mode: ProcessFrame returns once integration + raycast are queued (pose and status are final); the
        end-of-frame wait moves to the next ProcessFrame or to sync().  times() / stats() lag by one frame until then."""
        check(self.lib.xs_kinfu_set_deferred(self.h, 1 if on else 0), "set_deferred")

    def sync(self):
        check(self.lib.xs_kinfu_sync(self.h), "sync")

    @property
    def frame_id(self):
        return self.lib.xs_kinfu_frame_id(self.h)

    def getFrame(self):
        return self.frame_id

    @property
    def world2camera(self):
        """[(1+ncomp), 4, 4]: component 0 is the real matrix, the rest are the h-scaled derivative components."""
        out = np.zeros(((1 + self.ncomp) * 16,), np.float32)
        check(self.lib.xs_kinfu_get_world2camera(self.h, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out.reshape(1 + self.ncomp, 4, 4)

    @world2camera.setter
    def world2camera(self, value):
        """Only before the first frame: real part + every derivative component, [(1+ncomp), 4, 4]."""
        v = np.ascontiguousarray(value, np.float32).reshape(-1)
        if v.size != (1 + self.ncomp) * 16:
            raise ValueError("world2camera must hold (1+ncomp) 4x4 matrices")
        check(self.lib.xs_kinfu_set_world2camera(self.h, v.ctypes.data_as(C.POINTER(C.c_float))), "world2camera")

    def pose_c2w(self):
        """world2camera_record.back().inverse().real(), main.cpp:61"""
        out = np.zeros((16,), np.float32)
        check(self.lib.xs_kinfu_get_pose_c2w(self.h, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out.reshape(4, 4)

    def times(self):
        out = np.zeros((10,), np.float32)
        check(self.lib.xs_kinfu_get_times(self.h, out.ctypes.data_as(C.POINTER(C.c_float))))
        names = ("surface", "icp", "integrate", "raycast", "total")
        return {n: float(out[i]) for i, n in enumerate(names)}, {n: int(out[5 + i]) for i, n in enumerate(names)}

    def stats(self):
        out = (C.c_ulonglong * 4)()
        check(self.lib.xs_kinfu_get_stats(self.h, out))
        return [int(x) for x in out]

    def component_of(self, i, j=None):
        """Index (into world2camera[1:], volume planes, maps[1:]) of a Hessian batch's first-order component F_i, or of the
        second-order component S_ij of a listed pair."""
        if self.comps != 2:
            raise ValueError("component_of is defined for Hessian batches (comps = 2)")
        if j is None:
            return i
        return self.dirs + self.pairs.index((min(i, j), max(i, j)))

    def algorithmic_bytes(self):
        out = (C.c_double * 4)()
        check(self.lib.xs_kinfu_get_algorithmic_bytes(self.h, out))
        return dict(zip(("surface", "icp", "integrate", "raycast"), [float(x) for x in out]))

    def enable_icp_log(self, on=True):
        """The per-iteration normal equations stay on the device unless the log is enabled."""
        check(self.lib.xs_kinfu_enable_icp_log(self.h, 1 if on else 0))

    def icp_log(self, max_iters=16):
        buf = np.zeros((max_iters, 1 + self.ncomp, 42), np.float64)
        n = self.lib.xs_kinfu_take_icp_log(self.h, buf.ctypes.data_as(C.POINTER(C.c_double)), max_iters)
        return buf[:n]

    def map(self, which, level=0):
        """Device map as a torch tensor view copy.  which: depth, vmap_curr, nmap_curr, vmap_g_prev, nmap_g_prev."""
        import torch
        idx = {"depth": 0, "vmap_curr": 1, "nmap_curr": 2, "vmap_g_prev": 3, "nmap_g_prev": 4}[which]
        rows, cols, nc = C.c_int(), C.c_int(), C.c_int()
        p = self.lib.xs_kinfu_map(self.h, idx, level, C.byref(rows), C.byref(cols), C.byref(nc))
        shape = (rows.value, cols.value) if idx == 0 else (1 + nc.value, 3, rows.value, cols.value)
        return torch.as_tensor(_DeviceView(p, shape), device="cuda").clone()

    def volume_planes(self, comp=0):
        """Seam views of TsdfVolume::value/weight/grad (TsdfVolume.h:46-49) as dense [z, y, x] torch tensors."""
        import torch
        res = [int(self.config["tsdf_size_z"]), int(self.config["tsdf_size_y"]), int(self.config["tsdf_size_x"])]
        v = torch.empty(res, dtype=torch.float32, device="cuda")
        w = torch.empty(res, dtype=torch.int32, device="cuda")
        g = torch.empty(res, dtype=torch.float32, device="cuda") if self.ncomp else None
        vol = self.lib.xs_kinfu_volume(self.h)
        check(self.lib.xs_volume_export_planes(vol, comp, C.c_void_p(v.data_ptr()), C.c_void_p(w.data_ptr()),
                                               C.c_void_p(g.data_ptr()) if g is not None else None, None))
        torch.cuda.synchronize()
        return v, w, g

    def ExportPointCloud(self, max_buffer=1000000):
        """KinectFusionReconstruction.cpp:334-372.  Returns (points, normals) numpy [n, 3]."""
        import torch
        pts = torch.empty((max_buffer, 3), dtype=torch.float32, device="cuda")
        nrm = torch.empty((max_buffer, 3), dtype=torch.float32, device="cuda")
        n = self.lib.xs_extract_points(self.lib.xs_kinfu_volume(self.h), C.c_void_p(pts.data_ptr()),
                                       C.c_void_p(nrm.data_ptr()), max_buffer, None)
        check(n, "ExportPointCloud")
        return pts[:n].cpu().numpy(), nrm[:n].cpu().numpy()

    def set_comm(self, comm, record_floats):
        """Multi-GPU: attach the library's communicator (parallel.Comm); every later frame all-gathers the ranks' records."""
        self._comm = comm  # keep it alive
        check(self.lib.xs_kinfu_set_comm(self.h, comm.h if comm is not None else None, int(record_floats)), "set_comm")
        self._record_floats, self._world = int(record_floats), (comm.world if comm is not None else 1)

    def gathered_records(self, lag=0):
        """[world, record_floats] float32: the records of the last processed frame (lag = 0) or of the one before (lag = 1: its
        all-gather ran beside the last frame's kernels, so the read does not wait), gathered by the library over NCCL."""
        out = np.zeros((self._world, self._record_floats), np.float32)
        check(self.lib.xs_kinfu_get_gathered_records_lagged(self.h, int(lag), out.ctypes.data_as(C.POINTER(C.c_float))), "gathered_records")
        return out

    def pose_record_device_ptr(self):
        return self.lib.xs_kinfu_pose_record_device(self.h)

    def stream_ptr(self):
        """cudaStream_t (as an integer) all kernels of this object are launched on."""
        return self.lib.xs_kinfu_stream(self.h) or 0


# ------------------------------------------------------------------ outputs and synthetic input
def savePose(output_dir, frame_id, pose):
    """main.cpp:8-14 + IOHelper.cpp:21-32: frame-%06d.pose.txt, fixed, precision 7, trailing space."""
    os.makedirs(output_dir, exist_ok=True)
    path = os.path.join(output_dir, "frame-%06d.pose.txt" % frame_id)
    m = np.ascontiguousarray(pose, np.float32).reshape(16)
    check(_capi.load().xs_save_pose_txt(path.encode(), m.ctypes.data_as(C.POINTER(C.c_float))))
    return path


def exportPly(path, points, normals):
    """CPointCloud::exportPly, Visualization/src/CPointCloud.cpp:42-67"""
    p = np.ascontiguousarray(points, np.float32)
    n = np.ascontiguousarray(normals, np.float32)
    check(_capi.load().xs_export_ply(path.encode(), p.ctypes.data_as(C.POINTER(C.c_float)),
                                     n.ctypes.data_as(C.POINTER(C.c_float)), p.shape[0]))


def synth_pose(frame):
    out = np.zeros((16,), np.float32)
    check(_capi.load().xs_synth_pose(int(frame), out.ctypes.data_as(C.POINTER(C.c_float))))
    return out.reshape(4, 4)


def synth_depth(frame_or_pose, width=640, height=480, fx=481.20, fy=-480.00, cx=319.50, cy=239.50):
    """Synthetic ICL-NUIM-shaped depth frame (uint16 mm) for a frame index or an explicit c2w pose."""
    pose = synth_pose(frame_or_pose) if np.isscalar(frame_or_pose) else np.asarray(frame_or_pose, np.float32)
    pose = np.ascontiguousarray(pose, np.float32).reshape(16)
    out = np.zeros((height, width), np.uint16)
    check(_capi.load().xs_synth_depth(pose.ctypes.data_as(C.POINTER(C.c_float)), _capi.Intr(fx, fy, cx, cy), height, width,
                                      out.ctypes.data_as(C.POINTER(C.c_uint16))))
    return out
