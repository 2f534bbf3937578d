"""CPU checks of the dual-complex FP64 oracle (oracle/pyref.py OracleKinfu2): the witness tests/test_gpu_second_order_oracle.py
holds the product's second-order components against.
  - its value and first-order parts reproduce the existing FP64 complex oracle (different pose rounding and, for first order, the
    reference's Hermitian-LLT quirk apart);
  - the mixed second derivative is symmetric: a run in which parameter i rides on the imaginary unit and j on the dual part must
    return the same d2 / (d theta_i d theta_j) as the run with the roles swapped - the two runs share no derivative code path
    (complex arithmetic vs dual arithmetic), so agreement pins both."""
import ctypes
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
H_ = 1e-7


def _gen(a):
    G = np.zeros((4, 4))
    if a < 3:
        G[a, 3] = 1
    else:
        i, j = [(1, 2), (2, 0), (0, 1)][a - 3]
        G[i, j], G[j, i] = -1, 1
    return G


@pytest.fixture(scope="module")
def setup():
    from oracle import pyref
    import bench
    o = pyref.Oracle()
    W, Hh = 160, 120
    intr = (481.2 / 4, -480.0 / 4, 319.5 / 4, 239.5 / 4)
    cfg = dict(bench.REF_DEFAULTS)
    cfg.update(tsdf_size_x=64, tsdf_size_y=64, tsdf_size_z=64, tsdf_voxel_size=0.12, depth_width=W, depth_height=Hh, fx=intr[0], fy=intr[1],
               cx=intr[2], cy=intr[3])
    fp = ctypes.POINTER(ctypes.c_float)

    def depth(f):
        pose = np.zeros((16,), np.float32)
        o.lib.oracle_synth_pose(int(f), pose.ctypes.data_as(fp))
        out = np.zeros((Hh, W), np.uint16)
        o.lib.oracle_synth_depth(pose.ctypes.data_as(fp), ctypes.c_float(intr[0]), ctypes.c_float(intr[1]), ctypes.c_float(intr[2]),
                                 ctypes.c_float(intr[3]), Hh, W, out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)))
        return out

    return pyref, o, cfg, [depth(f) for f in range(3)]


def _run2(pyref, o, cfg, frames, i, j):
    Gi, Gj = _gen(i), _gen(j)
    k = pyref.OracleKinfu2(cfg, Gi, Gj, 0.5 * (Gi @ Gj + Gj @ Gi), H_, oracle=o)
    for d in frames:
        assert k.process_frame(d) == 1
    return k


@pytest.mark.parametrize("i,j", [(0, 4), (2, 3), (5, 1)])
def test_mixed_second_derivative_is_symmetric(setup, i, j):
    pyref, o, cfg, frames = setup
    a, b = _run2(pyref, o, cfg, frames, i, j), _run2(pyref, o, cfg, frames, j, i)
    s1, s2 = a.w2c.d.imag / H_, b.w2c.d.imag / H_
    assert np.abs(s1).max() > 1e-3
    assert np.abs(s1 - s2).max() <= 1e-5 * np.abs(s1).max()
    # and the first-order parts swap roles
    assert np.abs(a.w2c.m.imag / H_ - b.w2c.d.real).max() <= 1e-5 * np.abs(b.w2c.d.real).max()


def test_value_and_first_order_match_the_complex_oracle(setup):
    pyref, o, cfg, frames = setup
    k2 = _run2(pyref, o, cfg, frames, 0, 4)
    k1 = pyref.OracleKinfu(cfg, (H_ * _gen(4)).astype(np.float32), f64=True, oracle=o)
    for d in frames:
        assert k1.process_frame(d) == 1
    assert np.abs(k2.w2c.m.real - k1.w2c.real).max() <= 5e-6
    want = k1.w2c.imag / H_  # parameter 4 rides on the dual part of k2
    assert np.abs(k2.w2c.d.real - want).max() <= 1e-3 * np.abs(want).max()
