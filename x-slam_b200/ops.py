"""Operator-level host API: the reference's free functions (XKinectFusion/include/{Map,TsdfFusion,RayCaster,
ICP,ExtractPointCloud}.h) on torch CUDA tensors.

torch is used only for device memory and streams; every computation is a call through the C-ABI of
libxslam_b200.so.  Names and argument meaning follow the reference (bilateralFilter, pyrDown, createVMap,
createNMap, resizeVMap, resizeNMap, integrateTsdfVolume, raycast, estimateCombined,
ComputeLocalTsdf_hessian, ComputeLocalTsdf_loss, extractPoints/extractNormals); maps are packed SoA tensors
[(1+ncomp), 3, rows, cols] instead of interleaved pitched devComplex arrays.
"""
import ctypes as C

import numpy as np
import torch

from . import _capi
from ._capi import Intr, Pose, check


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not (t.is_cuda and t.is_contiguous()):
            raise ValueError("expected a contiguous CUDA tensor (libxslam_b200 has no CPU path)")


class PoseBatch:
    """A rigid transform with batched derivative components (host arrays); see xs_pose."""

    def __init__(self, R, t, dR=None, dt=None):
        self.R = np.ascontiguousarray(R, np.float32).reshape(3, 3)
        self.t = np.ascontiguousarray(t, np.float32).reshape(3)
        self.dR = None if dR is None else np.ascontiguousarray(dR, np.float32).reshape(-1, 9)
        self.dt = None if dt is None else np.ascontiguousarray(dt, np.float32).reshape(-1, 3)
        self.ncomp = 0 if self.dR is None else self.dR.shape[0]

    def c(self):
        p = Pose()
        p.R[:] = self.R.reshape(-1).tolist()
        p.t[:] = self.t.tolist()
        p.ncomp = self.ncomp
        if self.ncomp:
            p.dR = self.dR.ctypes.data_as(C.POINTER(C.c_float))
            p.dt = self.dt.ctypes.data_as(C.POINTER(C.c_float))
        return p


# ------------------------------------------------------------------ surface measurement (Map.h)
def bilateralFilter(depth_u16):
    """Map.h:16.  depth_u16: int16/uint16 CUDA tensor [rows, cols] (mm).  Returns float32 [rows, cols]."""
    _need_cuda(depth_u16)
    rows, cols = depth_u16.shape
    out = torch.empty((rows, cols), dtype=torch.float32, device=depth_u16.device)
    check(_capi.load().xs_bilateral_filter(_ptr(depth_u16), cols * 2, rows, cols, _ptr(out), _stream()), "bilateralFilter")
    return out


def pyrDown(src):
    """Map.h:22"""
    _need_cuda(src)
    rows, cols = src.shape
    out = torch.empty((rows // 2, cols // 2), dtype=torch.float32, device=src.device)
    check(_capi.load().xs_pyr_down(_ptr(src), rows, cols, _ptr(out), _stream()), "pyrDown")
    return out


def createVMap(intr, depth):
    """Map.h:29.  Returns float32 [3, rows, cols]."""
    _need_cuda(depth)
    rows, cols = depth.shape
    out = torch.empty((3, rows, cols), dtype=torch.float32, device=depth.device)
    check(_capi.load().xs_create_vmap(intr, _ptr(depth), rows, cols, _ptr(out), _stream()), "createVMap")
    return out


def createNMap(vmap):
    """Map.h:35"""
    _need_cuda(vmap)
    _, rows, cols = vmap.shape
    out = torch.empty_like(vmap)
    check(_capi.load().xs_create_nmap(_ptr(vmap), rows, cols, _ptr(out), _stream()), "createNMap")
    return out


def _resize(fn, m, comps):
    _need_cuda(m)
    nc1, _, rows, cols = m.shape
    if comps == 2:  # Hessian batch with all pairs: ncomp = n + n (n + 1) / 2
        dirs = int(round((np.sqrt(9 + 8 * (nc1 - 1)) - 3) / 2))
    else:
        dirs = (nc1 - 1) // comps
    out = torch.empty((nc1, 3, rows // 2, cols // 2), dtype=torch.float32, device=m.device)
    check(fn(_ptr(m), rows, cols, comps, dirs, _ptr(out), _stream()), "resizeMap")
    return out


def resizeVMap(m, comps=1):
    """Map.h:47.  m: [(1+ncomp), 3, rows, cols]"""
    return _resize(_capi.load().xs_resize_vmap, m, comps)


def resizeNMap(m, comps=1):
    """Map.h:54"""
    return _resize(_capi.load().xs_resize_nmap, m, comps)


# ------------------------------------------------------------------ TSDF volume (TsdfVolume.h, TsdfFusion.h, RayCaster.h)
class TsdfVolume:
    """TsdfVolume (TsdfVolume.h:18-60) over the brick-tiled device layout."""

    def __init__(self, resolution, voxel_size, thres_range, comps=1, dirs=0, pairs=None):
        """comps = 1 / 3: dirs first-order / bicomplex directions.  comps = 2: Hessian batch over dirs parameters with the
        second-order planes of `pairs` (default: all dirs (dirs + 1) / 2 pairs)."""
        self.lib = _capi.load()
        self.res = tuple(int(r) for r in resolution)
        self.comps, self.dirs = comps, dirs
        self.voxel_size = float(voxel_size)
        arr = (C.c_int * 3)(*self.res)
        if comps == 2:
            self.pairs = [(i, j) for i in range(dirs) for j in range(i, dirs)] if pairs is None else [tuple(p) for p in pairs]
            self.ncomp = dirs + len(self.pairs)
            pa = np.ascontiguousarray(np.asarray(self.pairs, np.int32).reshape(-1, 2))
            self.h = self.lib.xs_volume_create_hessian(arr, voxel_size, thres_range, dirs, len(self.pairs), pa.ctypes.data_as(C.POINTER(C.c_int)))
        else:
            self.ncomp = comps * dirs
            self.h = self.lib.xs_volume_create(arr, voxel_size, thres_range, comps, dirs)
        if not self.h:
            raise _capi.XsError("TsdfVolume: " + self.lib.xs_last_error().decode())

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.xs_volume_destroy(self.h)
            self.h = None

    def getTsdfTruncDist(self):
        return self.lib.xs_volume_trunc_dist(self.h)

    def reset(self):
        check(self.lib.xs_volume_reset(self.h, _stream()), "TsdfVolume.reset")

    def _dense(self, dtype):
        return torch.empty((self.res[2], self.res[1], self.res[0]), dtype=dtype, device="cuda")

    def value(self):
        v = self._dense(torch.float32)
        check(self.lib.xs_volume_export_planes(self.h, 0, _ptr(v), None, None, _stream()), "value")
        return v

    def weight(self):
        w = self._dense(torch.int32)
        check(self.lib.xs_volume_export_planes(self.h, 0, None, _ptr(w), None, _stream()), "weight")
        return w

    def grad(self, comp=0):
        g = self._dense(torch.float32)
        check(self.lib.xs_volume_export_planes(self.h, comp, None, None, _ptr(g), _stream()), "grad")
        return g

    def load(self, value=None, weight=None, grad=None, comp=0):
        _need_cuda(value, weight, grad)
        check(self.lib.xs_volume_import_planes(self.h, comp, _ptr(value), _ptr(weight), _ptr(grad), _stream()), "load")


def integrateTsdfVolume(depth_u16, intr, max_weight, volume, v2c, threshold=0.0):
    """TsdfFusion.h:40-45.  v2c: PoseBatch (Rv2c, tv2c).  Returns the number of updated voxels."""
    _need_cuda(depth_u16)
    rows, cols = depth_u16.shape
    stats = (C.c_ulonglong * 4)()
    p = v2c.c()
    check(volume.lib.xs_integrate(volume.h, _ptr(depth_u16), cols * 2, rows, cols, intr, max_weight, C.byref(p),
                                  threshold, stats, _stream()), "integrateTsdfVolume")
    return int(stats[0])


def raycast(intr, c2v, v2w, volume, rows, cols):
    """RayCaster.h:21-25.  Returns (vmap, nmap), each [(1+ncomp), 3, rows, cols] in the world frame."""
    vmap = torch.empty((1 + volume.ncomp, 3, rows, cols), dtype=torch.float32, device="cuda")
    nmap = torch.empty_like(vmap)
    pc, pw = c2v.c(), v2w.c()
    check(volume.lib.xs_raycast(volume.h, intr, C.byref(pc), C.byref(pw), rows, cols, _ptr(vmap), _ptr(nmap), _stream()),
          "raycast")
    return vmap, nmap


def ComputeLocalTsdf_hessian(depth_u16, intr, resolution, voxel_size, v2c, trunc, gt):
    """TsdfFusion.h:55-60.  v2c: PoseBatch with one bicomplex direction (ncomp = 3).  gt: float32 [z, y, x]."""
    _need_cuda(depth_u16, gt)
    rows, cols = depth_u16.shape
    res = (C.c_int * 3)(*[int(r) for r in resolution])
    out = (C.c_double * 4)()
    p = v2c.c()
    check(_capi.load().xs_tsdf_hessian(_ptr(depth_u16), cols * 2, rows, cols, intr, res, voxel_size, C.byref(p), trunc,
                                       _ptr(gt), out, _stream()), "ComputeLocalTsdf_hessian")
    return [out[i] for i in range(4)]


def ComputeLocalTsdf_hessian_batch(depth_u16, intr, resolution, voxel_size, v2c, trunc, gt):
    """ComputeLocalTsdf_hessian for dirs bicomplex directions at once (v2c: PoseBatch with ncomp = 3 * dirs); the ground-truth
    volume is streamed once per 8 directions.  Returns float64 [dirs, 4] = (sum loss, sum grad, sum hessian, count)."""
    _need_cuda(depth_u16, gt)
    rows, cols = depth_u16.shape
    res = (C.c_int * 3)(*[int(r) for r in resolution])
    p = v2c.c()
    dirs = v2c.ncomp // 3
    out = np.zeros((dirs, 4), np.float64)
    check(_capi.load().xs_tsdf_hessian_batch(_ptr(depth_u16), cols * 2, rows, cols, intr, res, voxel_size, C.byref(p), trunc,
                                             _ptr(gt), out.ctypes.data_as(C.POINTER(C.c_double)), _stream()),
          "ComputeLocalTsdf_hessian_batch")
    return out


def ComputeLocalTsdf_loss(depth_u16, intr, resolution, voxel_size, Rv2c, tv2c, trunc, gt):
    """TsdfFusion.h:48-52 (real-only volume loss).  Rv2c: float32 [3, 3] row-major, tv2c: float32 [3]; gt: float32 [z, y, x].
    Returns [sum loss, count]."""
    _need_cuda(depth_u16, gt)
    rows, cols = depth_u16.shape
    res = (C.c_int * 3)(*[int(r) for r in resolution])
    out = (C.c_double * 2)()
    R = (C.c_float * 9)(*[float(v) for v in np.asarray(Rv2c, np.float32).reshape(9)])
    t = (C.c_float * 3)(*[float(v) for v in np.asarray(tv2c, np.float32).reshape(3)])
    check(_capi.load().xs_tsdf_loss(_ptr(depth_u16), cols * 2, rows, cols, intr, res, voxel_size, R, t, trunc, _ptr(gt), out,
                                    _stream()), "ComputeLocalTsdf_loss")
    return [out[0], out[1]]


def extractPoints(volume, max_points=1000000, normals=True):
    """ExtractPointCloud.h:19-23 (extractPoints + extractNormals).  Returns (points [n,3], normals [n,3])."""
    pts = torch.empty((max_points, 3), dtype=torch.float32, device="cuda")
    nrm = torch.empty((max_points, 3), dtype=torch.float32, device="cuda") if normals else None
    n = volume.lib.xs_extract_points(volume.h, _ptr(pts), _ptr(nrm), max_points, _stream())
    check(n, "extractPoints")
    return pts[:n], (nrm[:n] if normals else None)


# ------------------------------------------------------------------ ICP (ICP.h)
def estimateCombined(curr, vmap_curr, nmap_curr, prev, intr, vmap_g_prev, nmap_g_prev, distThres, angleThres, comps=1):
    """ICP.h:24-31.  curr = PoseBatch(Rcurr, tcurr); prev = PoseBatch(Rprev_inv, tprev).
    Returns (A [(1+ncomp), 6, 6], b [(1+ncomp), 6]) as float64 numpy arrays (component 0 = real part)."""
    _need_cuda(vmap_curr, nmap_curr, vmap_g_prev, nmap_g_prev)
    _, rows, cols = vmap_curr.shape
    ncomp = vmap_g_prev.shape[0] - 1
    dirs = int(round((np.sqrt(9 + 8 * ncomp) - 3) / 2)) if comps == 2 else ncomp // comps  # comps = 2: all pairs of dirs parameters
    A = np.zeros((1 + ncomp, 36), np.float64)
    b = np.zeros((1 + ncomp, 6), np.float64)
    pc, pp = curr.c(), prev.c()
    check(_capi.load().xs_estimate_combined(C.byref(pc), _ptr(vmap_curr), _ptr(nmap_curr), C.byref(pp), intr,
                                            _ptr(vmap_g_prev), _ptr(nmap_g_prev), rows, cols, comps, dirs, distThres,
                                            angleThres, A.ctypes.data_as(C.POINTER(C.c_double)),
                                            b.ctypes.data_as(C.POINTER(C.c_double)), _stream()), "estimateCombined")
    # column-major 6x6 (symmetric, so the transpose is the same matrix)
    return A.reshape(1 + ncomp, 6, 6).transpose(0, 2, 1).copy(), b


def computeOptimizeMatrix(vmap_curr, nmap_curr, vmap_g_prev, nmap_g_prev, curr, prev, intr, distThres, angleThres):
    """ICP.h:34-40 (argument order as the reference).  Real parts only: vmap_g_prev / nmap_g_prev may carry derivative
    components, which are ignored.  Returns (count, jacobi [3, 4], hessian [12, 12]) - hessian[i1 * 4 + j1, i2 * 4 + j2] is
    the reference's hessian_host[i1][j1](i2, j2)."""
    _need_cuda(vmap_curr, nmap_curr, vmap_g_prev, nmap_g_prev)
    _, rows, cols = vmap_curr.shape
    J = np.zeros((12,), np.float64)
    H = np.zeros((144,), np.float64)
    pc, pp = curr.c(), prev.c()
    n = _capi.load().xs_compute_optimize_matrix(C.byref(pc), _ptr(vmap_curr), _ptr(nmap_curr), C.byref(pp), intr,
                                                _ptr(vmap_g_prev), _ptr(nmap_g_prev), rows, cols, distThres, angleThres,
                                                J.ctypes.data_as(C.POINTER(C.c_double)), H.ctypes.data_as(C.POINTER(C.c_double)),
                                                _stream())
    if n < 0:
        check(-1, "computeOptimizeMatrix")
    return int(n), J.reshape(3, 4), H.reshape(12, 12)


# ------------------------------------------------------------------ DeviceArray second-order complex (DoubleComplex)
DC_OPS = {"add": 0, "sub": 1, "mul": 2, "div": 3, "sqrt": 4, "exp": 5, "log": 6, "sin": 7, "cos": 8, "atan2": 9,
          "pow": 10, "atan": 11}


def dc_apply(op, a, b=None, p=0.0):
    """Elementwise bicomplex op on packed-SoA arrays a, b: float32 CUDA tensors [4, n]."""
    _need_cuda(a, b)
    n = a.shape[1]
    out = torch.empty_like(a)
    check(_capi.load().xs_dc_apply(DC_OPS[op], _ptr(a), _ptr(b), p, _ptr(out), n, _stream()), "dc_apply")
    return out


def dc_chain(t, h=1e-6):
    """test_CSFD main.cpp:194-205 evaluated at every t[i]; returns [4, n]."""
    _need_cuda(t)
    n = t.numel()
    out = torch.empty((4, n), dtype=torch.float32, device=t.device)
    check(_capi.load().xs_dc_chain(_ptr(t), h, _ptr(out), n, _stream()), "dc_chain")
    return out
