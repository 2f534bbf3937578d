"""ctypes access to the CHECKERS under oracle/ — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
  RefCuda   oracle/_ref/libxslam_ref.so   unmodified reference CUDA operators @sm_100a + restated orchestrator
  RefCsfd   oracle/_ref/libref_csfd.so    unmodified reference host bicomplex type (DoubleComplex.cpp)
  Oracle    oracle/_build/liboracle.so    this repo's CPU restatement of the hot path
Complex maps cross this boundary as float32 arrays with a trailing (re, im) axis.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_CUDA_PATH = os.path.join(_HERE, "_ref", "libxslam_ref.so")
REF_CSFD_PATH = os.path.join(_HERE, "_ref", "libref_csfd.so")
REF_TEST_CSFD = os.path.join(_HERE, "_ref", "ref_test_CSFD")
ORACLE_PATH = os.path.join(_HERE, "_build", "liboracle.so")


class XoConfig(C.Structure):
    _fields_ = [("res", C.c_int * 3), ("voxel_size", C.c_float), ("max_weight", C.c_int), ("thres_range", C.c_float),
                ("init_xyz", C.c_float * 3), ("r_deg", C.c_float * 3), ("width", C.c_int), ("height", C.c_int),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("num_levels", C.c_int),
                ("dist_thres", C.c_float), ("angle_thres_deg", C.c_float), ("bi_threshold", C.c_float),
                ("trunc_k", C.c_float)]


def make_xo_config(cfg):
    c = XoConfig()
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
    return c


def _f(a):
    return np.ascontiguousarray(a, np.float32)


def _p(a, t=C.c_float):
    return a.ctypes.data_as(C.POINTER(t))


def cpose(R, t):
    """complex 3x3 / 3-vector (numpy complex64) -> interleaved float arrays for the reference wrappers."""
    R = np.asarray(R, np.complex64).reshape(9)
    t = np.asarray(t, np.complex64).reshape(3)
    return _f(np.stack([R.real, R.imag], -1).reshape(-1)), _f(np.stack([t.real, t.imag], -1).reshape(-1))


class RefCuda:
    """The reference's own CUDA kernels (needs a GPU)."""

    def __init__(self):
        if not os.path.exists(REF_CUDA_PATH):
            raise RuntimeError("oracle/_ref/libxslam_ref.so not built (make -C oracle ref; needs /root/reference)")
        self.lib = C.CDLL(REF_CUDA_PATH)
        self.lib.ref_kinfu_create.restype = C.c_void_p
        self.lib.ref_kinfu_trunc.restype = C.c_float

    def bilateral(self, depth):
        d = np.ascontiguousarray(depth, np.uint16)
        out = np.zeros(d.shape + (2,), np.float32)
        self.lib.ref_bilateral(_p(d, C.c_uint16), d.shape[0], d.shape[1], _p(out))
        return out

    def pyrdown(self, src):
        s = _f(src)
        out = np.zeros((s.shape[0] // 2, s.shape[1] // 2, 2), np.float32)
        self.lib.ref_pyrdown(_p(s), s.shape[0], s.shape[1], _p(out))
        return out

    def vmap_nmap(self, depth_c, fx, fy, cx, cy):
        s = _f(depth_c)
        rows, cols = s.shape[:2]
        v = np.zeros((3, rows, cols, 2), np.float32)
        n = np.zeros((3, rows, cols, 2), np.float32)
        self.lib.ref_vmap_nmap(_p(s), rows, cols, C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy), _p(v), _p(n))
        return v, n

    def resize_map(self, m, normalize):
        s = _f(m)
        _, rows, cols, _ = s.shape
        out = np.zeros((3, rows // 2, cols // 2, 2), np.float32)
        self.lib.ref_resize_map(_p(s), rows, cols, int(normalize), _p(out))
        return out

    def integrate(self, depth, intr, max_weight, res, voxel, R, t, trunc, value, weight, grad, threshold=0.0):
        """value/weight/grad: dense [z, y, x] arrays, updated in place.  R, t: complex64.  Returns kernel ms."""
        d = np.ascontiguousarray(depth, np.uint16)
        Rf, tf = cpose(R, t)
        r = (C.c_int * 3)(*res)
        ms = C.c_float()
        self.lib.ref_integrate(_p(d, C.c_uint16), d.shape[0], d.shape[1], C.c_float(intr[0]), C.c_float(intr[1]),
                               C.c_float(intr[2]), C.c_float(intr[3]), int(max_weight), r, C.c_float(voxel), _p(Rf), _p(tf),
                               C.c_float(trunc), _p(value), _p(weight, C.c_int), _p(grad), C.c_float(threshold), C.byref(ms))
        return ms.value

    def raycast(self, intr, Rc2v, tc2v, Rv2w, tv2w, trunc, res, voxel, value, grad, rows, cols):
        a, b = cpose(Rc2v, tc2v)
        c, d = cpose(Rv2w, tv2w)
        r = (C.c_int * 3)(*res)
        v = np.zeros((3, rows, cols, 2), np.float32)
        n = np.zeros((3, rows, cols, 2), np.float32)
        ms = C.c_float()
        self.lib.ref_raycast(C.c_float(intr[0]), C.c_float(intr[1]), C.c_float(intr[2]), C.c_float(intr[3]), _p(a), _p(b),
                             _p(c), _p(d), C.c_float(trunc), r, C.c_float(voxel), _p(_f(value)), _p(_f(grad)), rows, cols,
                             _p(v), _p(n), C.byref(ms))
        return v, n, ms.value

    def estimate_combined(self, Rcurr, tcurr, vmap_curr, nmap_curr, Rprev_inv, tprev, intr, vmap_prev, nmap_prev,
                          dist_thres, angle_thres):
        a, b = cpose(Rcurr, tcurr)
        c, d = cpose(Rprev_inv, tprev)
        rows, cols = vmap_curr.shape[1:3]
        A = np.zeros((72,), np.float64)
        bb = np.zeros((12,), np.float64)
        ms = C.c_float()
        self.lib.ref_estimate_combined(_p(a), _p(b), _p(_f(vmap_curr)), _p(_f(nmap_curr)), _p(c), _p(d), C.c_float(intr[0]),
                                       C.c_float(intr[1]), C.c_float(intr[2]), C.c_float(intr[3]), _p(_f(vmap_prev)),
                                       _p(_f(nmap_prev)), rows, cols, C.c_float(dist_thres), C.c_float(angle_thres),
                                       _p(A, C.c_double), _p(bb, C.c_double), C.byref(ms))
        A = A.reshape(6, 6, 2)  # column-major entries (symmetric)
        bb = bb.reshape(6, 2)
        return (A[..., 0] + 1j * A[..., 1]).T.copy(), bb[:, 0] + 1j * bb[:, 1], ms.value

    def compute_optimize_matrix(self, Rcurr, tcurr, vmap_curr, nmap_curr, Rprev_inv, tprev, intr, vmap_prev, nmap_prev,
                                dist_thres, angle_thres):
        """ICP.cu:431 computeOptimizeMatrix.  Returns (J [3, 4], H [12, 12]) float32."""
        a, b = cpose(Rcurr, tcurr)
        c, d = cpose(Rprev_inv, tprev)
        rows, cols = vmap_curr.shape[1:3]
        J = np.zeros((12,), np.float32)
        H = np.zeros((144,), np.float32)
        self.lib.ref_compute_optimize_matrix(_p(a), _p(b), _p(_f(vmap_curr)), _p(_f(nmap_curr)), _p(c), _p(d),
                                             C.c_float(intr[0]), C.c_float(intr[1]), C.c_float(intr[2]), C.c_float(intr[3]),
                                             _p(_f(vmap_prev)), _p(_f(nmap_prev)), rows, cols, C.c_float(dist_thres),
                                             C.c_float(angle_thres), _p(J), _p(H))
        return J.reshape(3, 4), H.reshape(12, 12)

    def tsdf_hessian(self, depth, intr, res, voxel, R4, t4, trunc, gt):
        """R4: [9, 4], t4: [3, 4] bicomplex components (re.re, re.im, im.re, im.im)."""
        d = np.ascontiguousarray(depth, np.uint16)
        r = (C.c_int * 3)(*res)
        out = np.zeros((4,), np.float32)
        ms = C.c_float()
        self.lib.ref_tsdf_hessian(_p(d, C.c_uint16), d.shape[0], d.shape[1], C.c_float(intr[0]), C.c_float(intr[1]),
                                  C.c_float(intr[2]), C.c_float(intr[3]), r, C.c_float(voxel), _p(_f(R4)), _p(_f(t4)),
                                  C.c_float(trunc), _p(_f(gt)), _p(out), C.byref(ms))
        return out, ms.value

    def tsdf_loss(self, depth, intr, res, voxel, R, t, trunc, gt):
        """TsdfFusion.cu:409 ComputeLocalTsdf_loss.  R: [3, 3] float32 row-major, t: [3]."""
        d = np.ascontiguousarray(depth, np.uint16)
        r = (C.c_int * 3)(*res)
        out = np.zeros((2,), np.float32)
        ms = C.c_float()
        self.lib.ref_tsdf_loss(_p(d, C.c_uint16), d.shape[0], d.shape[1], C.c_float(intr[0]), C.c_float(intr[1]),
                               C.c_float(intr[2]), C.c_float(intr[3]), r, C.c_float(voxel), _p(_f(R)), _p(_f(t)),
                               C.c_float(trunc), _p(_f(gt)), _p(out), C.byref(ms))
        return out, ms.value

    def extract(self, res, voxel, value, weight, grad, max_points=1000000):
        """ExtractPointCloud.cu extractPoints + extractNormals as ExportPointCloud calls them.  Returns (points, normals)."""
        r = (C.c_int * 3)(*res)
        pts = np.zeros((max_points, 3), np.float32)
        nrm = np.zeros((max_points, 3), np.float32)
        self.lib.ref_extract.restype = C.c_long
        n = self.lib.ref_extract(r, C.c_float(voxel), _p(_f(value)), _p(np.ascontiguousarray(weight, np.int32), C.c_int), _p(_f(grad)),
                                 C.c_long(max_points), _p(pts), _p(nrm))
        return pts[:n], nrm[:n]

    def kinfu(self, cfg, seed_imag=None):
        return RefKinfu(self, cfg, seed_imag)


class RefKinfu:
    """Restated orchestrator driving the reference kernels, one perturbation direction per instance."""

    def __init__(self, ref, cfg, seed_imag=None):
        self.lib = ref.lib
        self.cfg = cfg
        c = make_xo_config(cfg)
        s = None
        if seed_imag is not None:
            s = _f(seed_imag).reshape(16)
        self.h = C.c_void_p(self.lib.ref_kinfu_create(C.byref(c), _p(s) if s is not None else None))
        self.W, self.H, self.L = int(cfg["depth_width"]), int(cfg["depth_height"]), int(cfg["num_levels"])
        self.res = (int(cfg["tsdf_size_x"]), int(cfg["tsdf_size_y"]), int(cfg["tsdf_size_z"]))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ref_kinfu_destroy(self.h)
            self.h = None

    def process_frame(self, depth):
        d = np.ascontiguousarray(depth, np.uint16)
        return self.lib.ref_kinfu_process_frame(self.h, _p(d, C.c_uint16))

    def pose(self):
        out = np.zeros((16, 2), np.float32)
        self.lib.ref_kinfu_get_pose(self.h, _p(out))
        return (out[:, 0] + 1j * out[:, 1]).reshape(4, 4)

    def times(self):
        out = np.zeros((5,), np.float32)
        self.lib.ref_kinfu_get_times(self.h, _p(out))
        return dict(zip(("surface", "icp", "integrate", "raycast", "total"), out.tolist()))

    def trunc(self):
        return self.lib.ref_kinfu_trunc(self.h)

    def map(self, which, level=0):
        idx = {"depth": 0, "vmap_curr": 1, "nmap_curr": 2, "vmap_g_prev": 3, "nmap_g_prev": 4}[which]
        r, c = self.H >> level, self.W >> level
        out = np.zeros(((r, c, 2) if idx == 0 else (3, r, c, 2)), np.float32)
        self.lib.ref_kinfu_get_map(self.h, idx, level, _p(out))
        return out

    def volume(self):
        shape = (self.res[2], self.res[1], self.res[0])
        v = np.zeros(shape, np.float32)
        w = np.zeros(shape, np.int32)
        g = np.zeros(shape, np.float32)
        self.lib.ref_kinfu_get_volume(self.h, _p(v), _p(w, C.c_int), _p(g))
        return v, w, g

    def icp_log(self, max_iters=16):
        buf = np.zeros((max_iters, 84), np.float64)
        n = self.lib.ref_kinfu_take_icp_log(self.h, _p(buf, C.c_double), max_iters)
        buf = buf[:n]
        A = buf[:, :72].reshape(n, 6, 6, 2)
        b = buf[:, 72:].reshape(n, 6, 2)
        return (A[..., 0] + 1j * A[..., 1]).transpose(0, 2, 1), b[..., 0] + 1j * b[..., 1]


class RefCsfd:
    """The reference's host DoubleComplex arithmetic (CPU)."""
    OPS = {"add": 0, "sub": 1, "mul": 2, "div": 3, "sqrt": 4, "exp": 5, "log": 6, "sin": 7, "cos": 8, "atan2": 9,
           "pow": 10, "atan": 11}

    def __init__(self):
        if not os.path.exists(REF_CSFD_PATH):
            raise RuntimeError("oracle/_ref/libref_csfd.so not built (make -C oracle ref; needs /root/reference)")
        self.lib = C.CDLL(REF_CSFD_PATH)
        self.lib.ref_dc_chain_bench.restype = C.c_double

    def apply(self, op, a, b=None, p=0.0):
        """a, b: [n, 4] float32 (AoS as in the reference: re.re, re.im, im.re, im.im)."""
        a = _f(a)
        out = np.zeros_like(a)
        bp = _p(_f(b)) if b is not None else None
        rc = self.lib.ref_dc_apply(self.OPS[op], _p(a), bp, C.c_float(p), _p(out), C.c_long(a.shape[0]))
        assert rc == 0
        return out

    def chain(self, t, h=1e-6):
        t = _f(t)
        out = np.zeros((t.shape[0], 4), np.float32)
        self.lib.ref_dc_chain(_p(t), C.c_float(h), _p(out), C.c_long(t.shape[0]))
        return out

    def chain_bench(self, t, h=1e-6, reps=1):
        t = _f(t)
        cs = C.c_double()
        rate = self.lib.ref_dc_chain_bench(_p(t), C.c_float(h), C.c_long(t.shape[0]), int(reps), C.byref(cs))
        return rate, cs.value


class Oracle:
    """This repo's CPU restatement (oracle/xslam_oracle.cpp, oracle/csfd_oracle.cpp).  f64=True selects the FP64
    arithmetic variant.  Complex maps are float32 arrays [3, rows, cols, 2]; poses are complex64."""

    def __init__(self):
        if not os.path.exists(ORACLE_PATH):
            raise RuntimeError("oracle/_build/liboracle.so not built (make -C oracle oracle)")
        self.lib = C.CDLL(ORACLE_PATH)
        self.lib.oracle_integrate.restype = C.c_long

    # ---- surface measurement
    def bilateral(self, depth):
        d = np.ascontiguousarray(depth, np.uint16)
        out = np.zeros(d.shape, np.float32)
        self.lib.oracle_bilateral(_p(d, C.c_uint16), d.shape[0], d.shape[1], _p(out))
        return out

    def pyrdown(self, src):
        s = _f(src)
        out = np.zeros((s.shape[0] // 2, s.shape[1] // 2), np.float32)
        self.lib.oracle_pyrdown(_p(s), s.shape[0], s.shape[1], _p(out))
        return out

    def vmap_nmap(self, depth, intr, f64=False):
        s = _f(depth)
        rows, cols = s.shape
        v = np.zeros((3, rows, cols, 2), np.float32)
        n = np.zeros((3, rows, cols, 2), np.float32)
        self.lib.oracle_vmap_nmap(_p(s), rows, cols, C.c_float(intr[0]), C.c_float(intr[1]), C.c_float(intr[2]),
                                  C.c_float(intr[3]), _p(v), _p(n), int(f64))
        return v, n

    def resize_map(self, m, normalize, f64=False):
        s = _f(m)
        _, rows, cols, _ = s.shape
        out = np.zeros((3, rows // 2, cols // 2, 2), np.float32)
        self.lib.oracle_resize_map(_p(s), rows, cols, int(normalize), _p(out), int(f64))
        return out

    # ---- volume
    def integrate(self, depth, intr, max_weight, res, voxel, R, t, trunc, value, weight, grad, threshold=0.0, f64=False,
                  z0=0, z1=None):
        d = np.ascontiguousarray(depth, np.uint16)
        Rf, tf = cpose(R, t)
        r = (C.c_int * 3)(*res)
        return self.lib.oracle_integrate(_p(d, C.c_uint16), d.shape[0], d.shape[1], C.c_float(intr[0]), C.c_float(intr[1]),
                                         C.c_float(intr[2]), C.c_float(intr[3]), int(max_weight), r, C.c_float(voxel), _p(Rf),
                                         _p(tf), C.c_float(trunc), _p(value), _p(weight, C.c_int), _p(grad),
                                         C.c_float(threshold), int(z0), int(res[2] if z1 is None else z1), int(f64))

    def raycast(self, intr, Rc2v, tc2v, Rv2w, tv2w, trunc, res, voxel, value, grad, rows, cols, f64=False, y0=0, y1=None):
        a, b = cpose(Rc2v, tc2v)
        c, d = cpose(Rv2w, tv2w)
        r = (C.c_int * 3)(*res)
        v = np.zeros((3, rows, cols, 2), np.float32)
        n = np.zeros((3, rows, cols, 2), np.float32)
        v[0, ..., 0] = np.nan
        n[0, ..., 0] = np.nan
        self.lib.oracle_raycast(C.c_float(intr[0]), C.c_float(intr[1]), C.c_float(intr[2]), C.c_float(intr[3]), _p(a), _p(b),
                                _p(c), _p(d), C.c_float(trunc), r, C.c_float(voxel), _p(_f(value)), _p(_f(grad)), rows, cols,
                                int(y0), int(rows if y1 is None else y1), _p(v), _p(n), int(f64))
        return v, n

    # ---- ICP
    def estimate_combined(self, Rcurr, tcurr, vmap_curr, nmap_curr, Rprev_inv, tprev, intr, vmap_prev, nmap_prev, dist_thres,
                          angle_thres, f64=False):
        a, b = cpose(Rcurr, tcurr)
        c, d = cpose(Rprev_inv, tprev)
        rows, cols = vmap_curr.shape[1:3]
        A = np.zeros((72,), np.float64)
        bb = np.zeros((12,), np.float64)
        self.lib.oracle_estimate_combined(_p(a), _p(b), _p(_f(vmap_curr)), _p(_f(nmap_curr)), _p(c), _p(d), C.c_float(intr[0]),
                                          C.c_float(intr[1]), C.c_float(intr[2]), C.c_float(intr[3]), _p(_f(vmap_prev)),
                                          _p(_f(nmap_prev)), rows, cols, C.c_float(dist_thres), C.c_float(angle_thres),
                                          _p(A, C.c_double), _p(bb, C.c_double), int(f64))
        A = A.reshape(6, 6, 2)
        bb = bb.reshape(6, 2)
        return (A[..., 0] + 1j * A[..., 1]).T.copy(), bb[:, 0] + 1j * bb[:, 1]

    def optimize_matrix(self, Rcurr, tcurr, vmap_curr, nmap_curr, Rprev_inv, tprev, intr, vmap_prev, nmap_prev, dist_thres,
                        angle_thres, f64=False):
        """computeOptimizeMatrix (ICP.cu:283-355, 431-489).  Returns (count, J [3, 4], H [12, 12]) float64."""
        a, b = cpose(Rcurr, tcurr)
        c, d = cpose(Rprev_inv, tprev)
        rows, cols = vmap_curr.shape[1:3]
        J = np.zeros((12,), np.float64)
        H = np.zeros((144,), np.float64)
        self.lib.oracle_optimize_matrix.restype = C.c_long
        n = self.lib.oracle_optimize_matrix(_p(a), _p(b), _p(_f(vmap_curr)), _p(_f(nmap_curr)), _p(c), _p(d), C.c_float(intr[0]),
                                            C.c_float(intr[1]), C.c_float(intr[2]), C.c_float(intr[3]), _p(_f(vmap_prev)),
                                            _p(_f(nmap_prev)), rows, cols, C.c_float(dist_thres), C.c_float(angle_thres),
                                            _p(J, C.c_double), _p(H, C.c_double), int(f64))
        return int(n), J.reshape(3, 4), H.reshape(12, 12)

    def tsdf_loss(self, depth, intr, res, voxel, R, t, trunc, gt, f64=False):
        """ComputeLocalTsdf_loss (TsdfFusion.cu:335-447).  Returns [sum loss, count]."""
        d = np.ascontiguousarray(depth, np.uint16)
        r = (C.c_int * 3)(*res)
        out = np.zeros((2,), np.float64)
        self.lib.oracle_tsdf_loss(_p(d, C.c_uint16), d.shape[0], d.shape[1], C.c_float(intr[0]), C.c_float(intr[1]),
                                  C.c_float(intr[2]), C.c_float(intr[3]), r, C.c_float(voxel), _p(_f(R)), _p(_f(t)),
                                  C.c_float(trunc), _p(_f(gt)), _p(out, C.c_double), int(f64))
        return out

    def pose_update(self, A, b, R, t):
        """One Gauss-Newton update (Hermitian-LLT solve as Eigen does); returns (ok, R, t)."""
        Ac = np.ascontiguousarray(np.stack([A.T.real, A.T.imag], -1), np.float64)  # column-major
        bc = np.ascontiguousarray(np.stack([b.real, b.imag], -1), np.float64)
        Rf, tf = cpose(R, t)
        ok = self.lib.oracle_pose_update(_p(Ac, C.c_double), _p(bc, C.c_double), _p(Rf), _p(tf))
        Rf = Rf.reshape(9, 2)
        tf = tf.reshape(3, 2)
        return ok, (Rf[:, 0] + 1j * Rf[:, 1]).reshape(3, 3).astype(np.complex64), (tf[:, 0] + 1j * tf[:, 1]).astype(np.complex64)

    def set_intrinsic_imag(self, fx0, dfx=0.0, dfy=0.0, dcx=0.0, dcy=0.0):
        """Complex intrinsics for the following calls (an extension: see oracle/xslam_oracle.cpp complex_intr)."""
        self.lib.oracle_set_intrinsic_imag(C.c_float(fx0), C.c_float(dfx), C.c_float(dfy), C.c_float(dcx), C.c_float(dcy))

    def llt_solve6(self, A, b):
        """The restated Eigen LLT (lower, Hermitian semantics) solve alone; A: [6, 6] complex, b: [6] complex."""
        Ac = np.ascontiguousarray(np.stack([A.T.real, A.T.imag], -1), np.float64)  # column-major
        bc = np.ascontiguousarray(np.stack([b.real, b.imag], -1), np.float64)
        x = np.zeros((6, 2), np.float64)
        ok = self.lib.oracle_llt_solve6(_p(Ac, C.c_double), _p(bc, C.c_double), _p(x, C.c_double))
        return ok, x[:, 0] + 1j * x[:, 1]

    def _m(self, fn, M, n):
        a = np.asarray(M, np.complex64).reshape(-1)
        buf = _f(np.stack([a.real, a.imag], -1).reshape(-1))
        out = np.zeros_like(buf)
        fn(_p(buf), _p(out))
        out = out.reshape(-1, 2)
        return (out[:, 0] + 1j * out[:, 1]).reshape(n, n).astype(np.complex64)

    def inverse4(self, M):
        return self._m(self.lib.oracle_inverse4, M, 4)

    def inverse3(self, M):
        return self._m(self.lib.oracle_inverse3, M, 3)

    def mul4(self, A, B):
        a = np.asarray(A, np.complex64).reshape(-1)
        b = np.asarray(B, np.complex64).reshape(-1)
        fa, fb = _f(np.stack([a.real, a.imag], -1).reshape(-1)), _f(np.stack([b.real, b.imag], -1).reshape(-1))
        out = np.zeros_like(fa)
        self.lib.oracle_mul4(_p(fa), _p(fb), _p(out))
        out = out.reshape(-1, 2)
        return (out[:, 0] + 1j * out[:, 1]).reshape(4, 4).astype(np.complex64)

    # ---- DoubleComplex arrays ([n, 4])
    def dc_apply(self, op, a, b=None, p=0.0, f64=False):
        dt = np.float64 if f64 else np.float32
        ct = C.c_double if f64 else C.c_float
        a = np.ascontiguousarray(a, dt)
        out = np.zeros_like(a)
        bp = _p(np.ascontiguousarray(b, dt), ct) if b is not None else None
        fn = self.lib.oracle_dc_apply_f64 if f64 else self.lib.oracle_dc_apply_f32
        rc = fn(RefCsfd.OPS[op], _p(a, ct), bp, ct(p), _p(out, ct), C.c_long(a.shape[0]))
        assert rc == 0
        return out

    def dc_chain(self, t, h=1e-6, f64=False):
        dt = np.float64 if f64 else np.float32
        ct = C.c_double if f64 else C.c_float
        t = np.ascontiguousarray(t, dt)
        out = np.zeros((t.shape[0], 4), dt)
        (self.lib.oracle_dc_chain_f64 if f64 else self.lib.oracle_dc_chain_f32)(_p(t, ct), ct(h), _p(out, ct), C.c_long(t.shape[0]))
        return out


class OracleKinfu:
    """CPU port of the frame loop (KinectFusionReconstruction.cpp:147-332) on the oracle stages: one perturbation
    direction per instance, as the reference.  z_slab / row_band bound the volume sweep and the raycast for the
    timed cpu_baseline sample (the default processes everything)."""

    def __init__(self, cfg, seed_imag=None, f64=False, oracle=None):
        self.o = oracle or Oracle()
        self.cfg = cfg
        self.f64 = f64
        self.W, self.H, self.L = int(cfg["depth_width"]), int(cfg["depth_height"]), int(cfg["num_levels"])
        self.res = (int(cfg["tsdf_size_x"]), int(cfg["tsdf_size_y"]), int(cfg["tsdf_size_z"]))
        self.voxel = float(np.float32(cfg["tsdf_voxel_size"]))
        self.trunc = float(max(np.float32(self.voxel) * np.float32(cfg["thres_range"]), np.float32(2.1) * np.float32(self.voxel)))
        self.intr = tuple(float(np.float32(cfg[k])) for k in ("fx", "fy", "cx", "cy"))
        self.w2c = np.eye(4, dtype=np.complex64)
        if seed_imag is not None:
            self.w2c = (self.w2c + 1j * np.asarray(seed_imag, np.float32).reshape(4, 4)).astype(np.complex64)
        self.record = [self.w2c]
        self.w2v = np.eye(4, dtype=np.complex64)
        self.w2v[:3, 3] = [cfg["init_x"], cfg["init_y"], cfg["init_z"]]
        self.angle_thres = float(np.float32(np.sin(np.float32(cfg["angleThres"]) / np.float32(180.0) * np.pi)))
        self.dist_thres = float(cfg["distThres"])
        shape = (self.res[2], self.res[1], self.res[0])
        self.value = np.zeros(shape, np.float32)
        self.weight = np.zeros(shape, np.int32)
        self.grad = np.zeros(shape, np.float32)
        self.vprev = [None] * self.L
        self.nprev = [None] * self.L
        self.frame_id = 0
        self.iters = [5, 4, 3]
        self.icp_log = []

    def level_intr(self, i):
        d = np.float32(1 << i)
        return tuple(float(np.float32(v) / d) for v in self.intr)

    def process_frame(self, depth, z_slab=None, row_band=None):
        o = self.o
        # SurfaceMeasure
        lev = [o.bilateral(depth)]
        for i in range(1, self.L):
            lev.append(o.pyrdown(lev[-1]))
        vc, nc = [], []
        for i in range(self.L):
            v, n = o.vmap_nmap(lev[i], self.level_intr(i), self.f64)
            vc.append(v)
            nc.append(n)
        # PoseEstimate
        self.icp_log = []
        if self.frame_id > 0:
            c2w_prev = o.inverse4(self.record[-1])
            Rprev, tprev = c2w_prev[:3, :3], c2w_prev[:3, 3]
            Rprev_inv = o.inverse3(Rprev)
            Rcurr, tcurr = Rprev.copy(), tprev.copy()
            for level in range(self.L - 1, -1, -1):
                for _ in range(self.iters[level]):
                    A, b = o.estimate_combined(Rcurr, tcurr, vc[level], nc[level], Rprev_inv, tprev, self.level_intr(level),
                                               self.vprev[level], self.nprev[level], self.dist_thres, self.angle_thres, self.f64)
                    self.icp_log.append((A, b))
                    ok, Rcurr, tcurr = o.pose_update(A, b, Rcurr, tcurr)
                    if not ok:
                        return 0
            c2w = np.eye(4, dtype=np.complex64)
            c2w[:3, :3], c2w[:3, 3] = Rcurr, tcurr
            self.w2c = o.inverse4(c2w)
            self.record.append(self.w2c)
        # IntegrateFrame
        c2w = o.inverse4(self.record[-1])
        c2v = o.mul4(self.w2v, c2w)
        v2c = o.inverse4(c2v)
        z0, z1 = (0, self.res[2]) if z_slab is None else z_slab
        self.updated = o.integrate(depth, self.intr, int(self.cfg["max_integration_weight"]), self.res, self.voxel, v2c[:3, :3],
                                   v2c[:3, 3], self.trunc, self.value, self.weight, self.grad,
                                   float(self.cfg["biInterpolate_threshold"]), self.f64, z0, z1)
        # raycast + pyramid
        v2w = o.inverse4(self.w2v)
        y0, y1 = (0, self.H) if row_band is None else row_band
        v, n = o.raycast(self.intr, c2v[:3, :3], c2v[:3, 3], v2w[:3, :3], v2w[:3, 3], self.trunc, self.res, self.voxel, self.value,
                         self.grad, self.H, self.W, self.f64, y0, y1)
        self.vprev[0], self.nprev[0] = v, n
        for i in range(1, self.L):
            self.vprev[i] = o.resize_map(self.vprev[i - 1], False, self.f64)
            self.nprev[i] = o.resize_map(self.nprev[i - 1], True, self.f64)
        self.frame_id += 1
        return 1


# ------------------------------------------------------------------ second order in FP64: the dual-complex frame loop
class DC:
    """A matrix of dual-complex numbers as a pair of complex128 arrays: m = (re.v + i im.v), d = (re.d + i im.d), i.e. the
    value and its exact derivative along one more parameter theta_j (oracle_types.h Dual).  The imaginary unit carries
    h d/d theta_i as everywhere in the reference, so d.imag / h is the mixed second derivative."""

    def __init__(self, m, d=None):
        self.m = np.asarray(m, np.complex128)
        self.d = np.zeros_like(self.m) if d is None else np.asarray(d, np.complex128)

    def __matmul__(self, o):
        return DC(self.m @ o.m, self.d @ o.m + self.m @ o.d)

    def __add__(self, o):
        return DC(self.m + o.m, self.d + o.d)

    def inv(self):
        mi = np.linalg.inv(self.m)
        return DC(mi, -mi @ self.d @ mi)

    def sub(self, rows, cols):
        return DC(self.m[rows, cols], self.d[rows, cols])

    def floats(self):
        """[2][n][2] float32: primary (re, im) pairs followed by the dual-part pairs (the layout the Dual stages read)."""
        f = lambda a: np.stack([a.reshape(-1).real, a.reshape(-1).imag], -1)
        return np.ascontiguousarray(np.stack([f(self.m), f(self.d)]), np.float32)


def _dc_axis(angle_m, angle_d, axis):
    """AngleAxis(angle, unit axis).toRotationMatrix() (host_algebra.h angle_axis_matrix) with its derivative."""
    c, s = np.cos(angle_m), np.sin(angle_m)
    dc, ds = -s * angle_d, c * angle_d
    i, j = [(1, 2), (2, 0), (0, 1)][axis]
    R, dR = np.eye(3, dtype=np.complex128), np.zeros((3, 3), np.complex128)
    R[i, i] = R[j, j] = c
    R[i, j], R[j, i] = -s, s
    dR[i, i] = dR[j, j] = dc
    dR[i, j], dR[j, i] = -ds, ds
    return DC(R, dR)


class OracleKinfu2:
    """The frame loop of OracleKinfu in dual-complex FP64 arithmetic (the stages instantiated with T = Dual, the pose algebra in
    numpy): one PAIR of perturbation parameters (i, j) per instance.  seed_i, seed_j, seed_ij: 4x4 real matrices G_i, G_j,
    (G_i G_j + G_j G_i) / 2 - the same seeds the product's Hessian batch takes, without their h factors.  After a frame
      w2c.m.imag / h  = d w2c / d theta_i,   w2c.d.real = d w2c / d theta_j,   w2c.d.imag / h = d2 w2c / (d theta_i d theta_j).
    The Gauss-Newton solve is the plain complex(-symmetric) one, i.e. the product's XS_SOLVE_ANALYTIC semantics: second order has
    no reference semantics to follow (the reference has no DCSFD frame loop)."""

    def __init__(self, cfg, seed_i, seed_j, seed_ij, h=1e-7, oracle=None):
        self.o = oracle or Oracle()
        self.cfg, self.h = cfg, h
        self.W, self.H, self.L = int(cfg["depth_width"]), int(cfg["depth_height"]), int(cfg["num_levels"])
        self.res = (int(cfg["tsdf_size_x"]), int(cfg["tsdf_size_y"]), int(cfg["tsdf_size_z"]))
        self.voxel = float(np.float32(cfg["tsdf_voxel_size"]))
        self.trunc = float(max(np.float32(self.voxel) * np.float32(cfg["thres_range"]), np.float32(2.1) * np.float32(self.voxel)))
        self.intr = tuple(float(np.float32(cfg[k])) for k in ("fx", "fy", "cx", "cy"))
        G = lambda a: np.asarray(a, np.float64).reshape(4, 4)
        self.w2c = DC(np.eye(4) + 1j * h * G(seed_i), G(seed_j) + 1j * h * G(seed_ij))
        self.record = [self.w2c]
        w2v = np.eye(4)
        w2v[:3, 3] = [cfg["init_x"], cfg["init_y"], cfg["init_z"]]
        self.w2v = DC(w2v)
        self.angle_thres = float(np.float32(np.sin(np.float32(cfg["angleThres"]) / np.float32(180.0) * np.pi)))
        self.dist_thres = float(cfg["distThres"])
        shape = (2, self.res[2], self.res[1], self.res[0])  # [0] values, [1] dual parts
        self.value = np.zeros(shape, np.float32)
        self.grad = np.zeros(shape, np.float32)
        self.weight = np.zeros(shape[1:], np.int32)
        self.vprev, self.nprev = [None] * self.L, [None] * self.L
        self.frame_id = 0
        self.iters = [5, 4, 3]

    def level_intr(self, i):
        d = np.float32(1 << i)
        return tuple(float(np.float32(v) / d) for v in self.intr)

    def _maps2(self, rows, cols):
        m = np.zeros((2, 3, rows, cols, 2), np.float32)
        m[0, 0, ..., 0] = np.nan
        return m

    def _estimate(self, Rc, tc, vc, nc, Rpi, tp, intr, vp, npv):
        lib = self.o.lib
        rows, cols = vc.shape[2:4]
        A = np.zeros((2, 72), np.float64)
        b = np.zeros((2, 12), np.float64)
        lib.oracle_estimate_combined(_p(Rc.floats()), _p(tc.floats()), _p(vc), _p(nc), _p(Rpi.floats()), _p(tp.floats()),
                                     C.c_float(intr[0]), C.c_float(intr[1]), C.c_float(intr[2]), C.c_float(intr[3]), _p(vp), _p(npv),
                                     rows, cols, C.c_float(self.dist_thres), C.c_float(self.angle_thres), _p(A, C.c_double),
                                     _p(b, C.c_double), 2)
        cA = lambda a: (a.reshape(6, 6, 2)[..., 0] + 1j * a.reshape(6, 6, 2)[..., 1]).T.copy()
        cb = lambda a: a.reshape(6, 2)[:, 0] + 1j * a.reshape(6, 2)[:, 1]
        return DC(cA(A[0]), cA(A[1])), DC(cb(b[0]), cb(b[1]))

    def process_frame(self, depth):
        o, lib = self.o, self.o.lib
        lev = [o.bilateral(depth)]
        for i in range(1, self.L):
            lev.append(o.pyrdown(lev[-1]))
        vc, nc = [], []
        for i in range(self.L):  # the current-frame maps do not depend on the pose: zero dual parts
            v, n = o.vmap_nmap(lev[i], self.level_intr(i), True)
            vc.append(np.ascontiguousarray(np.stack([v, np.zeros_like(v)])))
            nc.append(np.ascontiguousarray(np.stack([n, np.zeros_like(n)])))
        if self.frame_id > 0:
            c2w_prev = self.record[-1].inv()
            Rprev, tprev = c2w_prev.sub(slice(0, 3), slice(0, 3)), c2w_prev.sub(slice(0, 3), 3)
            Rprev_inv = Rprev.inv()
            Rcurr, tcurr = DC(Rprev.m.copy(), Rprev.d.copy()), DC(tprev.m.copy(), tprev.d.copy())
            for level in range(self.L - 1, -1, -1):
                for _ in range(self.iters[level]):
                    A, b = self._estimate(Rcurr, tcurr, vc[level], nc[level], Rprev_inv, tprev, self.level_intr(level),
                                          self.vprev[level], self.nprev[level])
                    x = np.linalg.solve(A.m, b.m)
                    dx = np.linalg.solve(A.m, b.d - A.d @ x)
                    Rinc = _dc_axis(x[2], dx[2], 2) @ _dc_axis(x[1], dx[1], 1) @ _dc_axis(x[0], dx[0], 0)
                    tn = DC(Rinc.m @ tcurr.m + x[3:], Rinc.d @ tcurr.m + Rinc.m @ tcurr.d + dx[3:])
                    Rcurr, tcurr = Rinc @ Rcurr, tn
            c2w = DC(np.eye(4))
            c2w.m[:3, :3], c2w.m[:3, 3], c2w.d[:3, :3], c2w.d[:3, 3] = Rcurr.m, tcurr.m, Rcurr.d, tcurr.d
            self.w2c = c2w.inv()
            self.record.append(self.w2c)
        c2w = self.record[-1].inv()
        c2v = self.w2v @ c2w
        v2c = c2v.inv()
        r = (C.c_int * 3)(*self.res)
        d = np.ascontiguousarray(depth, np.uint16)
        R3, t3 = lambda M: M.sub(slice(0, 3), slice(0, 3)).floats(), lambda M: M.sub(slice(0, 3), 3).floats()
        lib.oracle_integrate(_p(d, C.c_uint16), d.shape[0], d.shape[1], C.c_float(self.intr[0]), C.c_float(self.intr[1]),
                             C.c_float(self.intr[2]), C.c_float(self.intr[3]), int(self.cfg["max_integration_weight"]), r,
                             C.c_float(self.voxel), _p(R3(v2c)), _p(t3(v2c)), C.c_float(self.trunc), _p(self.value),
                             _p(self.weight, C.c_int), _p(self.grad), C.c_float(float(self.cfg["biInterpolate_threshold"])), 0,
                             int(self.res[2]), 2)
        v2w = self.w2v.inv()
        v, n = self._maps2(self.H, self.W), self._maps2(self.H, self.W)
        lib.oracle_raycast(C.c_float(self.intr[0]), C.c_float(self.intr[1]), C.c_float(self.intr[2]), C.c_float(self.intr[3]),
                           _p(R3(c2v)), _p(t3(c2v)), _p(R3(v2w)), _p(t3(v2w)), C.c_float(self.trunc), r, C.c_float(self.voxel),
                           _p(self.value), _p(self.grad), self.H, self.W, 0, self.H, _p(v), _p(n), 2)
        self.vprev[0], self.nprev[0] = v, n
        for i in range(1, self.L):
            rows, cols = self.vprev[i - 1].shape[2:4]
            for src, dst, norm in ((self.vprev, self.vprev, 0), (self.nprev, self.nprev, 1)):
                out = np.zeros((2, 3, rows // 2, cols // 2, 2), np.float32)
                lib.oracle_resize_map(_p(src[i - 1]), rows, cols, norm, _p(out), 2)
                dst[i] = out
        self.frame_id += 1
        return 1
