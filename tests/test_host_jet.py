"""CPU test of the host jet algebra (x-slam_b200/csrc/host_jet.h): the 4x4 / 3x3 cofactor inverses, products and axis
rotations the frame loop evaluates on the host between ICP and integration (KinectFusionReconstruction.cpp:167-173,231,
248-258,305-320), for first-order (CSFD) and second-order (DCSFD: eps1, eps2, eps1eps2) components, against the analytic
matrix derivatives in float64:

    d(A^-1)      = -A^-1 dA A^-1
    d12(A^-1)    = -A^-1 d12A A^-1 + A^-1 d1A A^-1 d2A A^-1 + A^-1 d2A A^-1 d1A A^-1
    d12(A B)     = d12A B + A d12B + d1A d2B + d2A d1B
    d12 R(theta) = R'(theta) d12theta + R''(theta) d1theta d2theta

Tolerance: FP32 evaluation, 2e-5 of the largest entry of each component."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("hj") / "host_jet_harness")
    subprocess.run(["g++", "-std=c++17", "-O3", "-o", exe, os.path.join(ROOT, "tests", "harness", "host_jet_harness.cpp")],
                   check=True)
    return exe


def rigid(rng):
    q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    T = np.eye(4)
    T[:3, :3] = q
    T[:3, 3] = rng.standard_normal(3) * 2
    return T


def rot(axis, t, order=0):
    """rotation about a coordinate axis and its first / second derivative with respect to the angle"""
    s, c = {0: (np.sin(t), np.cos(t)), 1: (np.cos(t), -np.sin(t)), 2: (-np.sin(t), -np.cos(t))}[order]
    R = np.zeros((3, 3))
    for i in range(3):
        R[i, i] = c if i != axis else (1.0 if order == 0 else 0.0)
    a, b = (axis + 1) % 3, (axis + 2) % 3
    R[a, b], R[b, a] = -s, s
    return R


def run(harness, comps, dirs, A, B, angle):
    txt = "%d %d\n" % (comps, dirs) + "\n".join("%.9g" % x for x in np.concatenate([A.ravel(), B.ravel(), angle.ravel()]))
    out = subprocess.run([harness], input=txt, capture_output=True, text=True, check=True).stdout.split()
    v = np.array([float(x) for x in out])
    n = 1 + comps * dirs
    o = 0
    res = []
    for sz in (16, 16, 9, 9):
        side = 4 if sz == 16 else 3
        res.append(v[o:o + n * sz].reshape(n, side, side))
        o += n * sz
    assert o == v.size
    return res


def close(a, b, what):
    scale = max(np.abs(b).max(), 1e-30)
    assert np.abs(a - b).max() <= 2e-5 * scale, "%s: %g of %g" % (what, np.abs(a - b).max(), scale)


@pytest.mark.parametrize("comps,dirs", [(1, 6), (3, 4), (3, 55)])
def test_host_pose_algebra_matches_analytic_derivatives(harness, comps, dirs):
    rng = np.random.default_rng(11 + comps + dirs)
    n = comps * dirs
    A = np.concatenate([rigid(rng)[None], rng.standard_normal((n, 4, 4))]).astype(np.float32).astype(np.float64)
    B = np.concatenate([rigid(rng)[None], rng.standard_normal((n, 4, 4))]).astype(np.float32).astype(np.float64)
    A[0, 3], B[0, 3] = (0, 0, 0, 1), (0, 0, 0, 1)
    angle = np.concatenate([[0.37], rng.standard_normal(n)]).astype(np.float32).astype(np.float64)
    inv, prod, inv3, rzyx = run(harness, comps, dirs, A, B, angle)

    def groups():
        if comps == 1:
            for q in range(dirs):
                yield (1 + q, None, None)
        else:
            for k in range(dirs):
                yield (1 + 3 * k, 2 + 3 * k, 3 + 3 * k)

    for M, got, tag in ((A, inv, "inverse4"), (A[:, :3, :3], inv3, "inverse3")):
        Mi = np.linalg.inv(M[0])
        close(got[0], Mi, tag + " real")
        for i1, i2, i12 in groups():
            close(got[i1], -Mi @ M[i1] @ Mi, tag + " first order")
            if i2 is not None:
                close(got[i2], -Mi @ M[i2] @ Mi, tag + " first order (eps2)")
                close(got[i12], -Mi @ M[i12] @ Mi + Mi @ M[i1] @ Mi @ M[i2] @ Mi + Mi @ M[i2] @ Mi @ M[i1] @ Mi, tag + " second order")
    close(prod[0], A[0] @ B[0], "product real")
    for i1, i2, i12 in groups():
        close(prod[i1], A[i1] @ B[0] + A[0] @ B[i1], "product first order")
        if i2 is not None:
            close(prod[i12], A[i12] @ B[0] + A[0] @ B[i12] + A[i1] @ B[i2] + A[i2] @ B[i1], "product second order")
    # Rinc = Rz * Ry * Rx of one batched angle (KinectFusionReconstruction.cpp:215-218)
    t = angle[0]
    R = [rot(ax, t) for ax in (2, 1, 0)]
    R1 = [rot(ax, t, 1) for ax in (2, 1, 0)]
    R2 = [rot(ax, t, 2) for ax in (2, 1, 0)]
    f0 = R[0] @ R[1] @ R[2]
    f1 = R1[0] @ R[1] @ R[2] + R[0] @ R1[1] @ R[2] + R[0] @ R[1] @ R1[2]
    f2 = (R2[0] @ R[1] @ R[2] + R[0] @ R2[1] @ R[2] + R[0] @ R[1] @ R2[2] +
          2 * (R1[0] @ R1[1] @ R[2] + R1[0] @ R[1] @ R1[2] + R[0] @ R1[1] @ R1[2]))
    close(rzyx[0], f0, "rotation real")
    for i1, i2, i12 in groups():
        close(rzyx[i1], f1 * angle[i1], "rotation first order")
        if i2 is not None:
            close(rzyx[i12], f1 * angle[i12] + f2 * angle[i1] * angle[i2], "rotation second order")


def test_pose_chain_matches_reference_complex_semantics(tmp_path):
    """The product's host jets against the oracle's restatement of the reference's Eigen / std::complex<float> pose chain
    (world2camera -> c2w -> Rprev_inv, c2v, v2c) with an h-scaled se(3) perturbation in the imaginary part: real parts are
    bit-identical (>= 97 % of the entries; the rest are structural zeros where the complex products leave terms of size
    h^2 ~ 1e-14 times rounding, far below one ulp of the matrix), derivative parts within 1e-5 (measured 9e-8)."""
    import sys
    sys.path.insert(0, ROOT)
    import xslam_b200 as xs
    exe = str(tmp_path / "host_jet_vs_oracle")
    subprocess.run(["g++", "-std=c++17", "-O2", "-o", exe, os.path.join(ROOT, "tests", "harness", "host_jet_vs_oracle.cpp")],
                   check=True)
    G = xs.se3_generators()
    w2v = np.eye(4)
    w2v[:3, 3] = 3.2
    worst_real, worst_deriv, identical, total = 0.0, 0.0, 0, 0
    for frame in (1, 7, 40):
        w2c = np.linalg.inv(xs.synth_pose(frame).astype(np.float64))
        for q in range(6):
            seed = 1e-7 * (G[q] @ w2c)  # derivative of exp(h G_q) * world2camera
            txt = "\n".join("%.9g" % x for x in np.concatenate([w2c.ravel(), seed.ravel(), w2v.ravel()]))
            out = subprocess.run([exe], input=txt, capture_output=True, text=True, check=True).stdout.split()
            v = np.array([float(x) for x in out]).reshape(-1, 4)
            rj, ro, dj, do = v[:, 0].astype(np.float32), v[:, 1].astype(np.float32), v[:, 2], v[:, 3]
            worst_real = max(worst_real, float(np.abs(rj.astype(np.float64) - ro).max() / np.abs(ro).max()))
            identical += int((rj == ro).sum())
            total += rj.size
            scale = np.abs(do).max()
            assert scale > 1e-9
            worst_deriv = max(worst_deriv, float(np.abs(dj - do).max() / scale))
    print("[host jets vs oracle] real: %d of %d entries bit-identical, worst |diff| / max = %.3g; derivative rel %.3g"
          % (identical, total, worst_real, worst_deriv))
    assert identical >= 0.97 * total
    assert worst_real <= 2.0 ** -23, worst_real  # one ulp of the largest entry
    assert worst_deriv <= 1e-5, worst_deriv


def test_hessian_batch_equals_dcsfd_list_on_the_host(tmp_path):
    """The host jets of a Hessian batch (comps = 2: first-order components stored once, one second-order component per listed
    pair) against the DCSFD list of the same pairs on a pose chain of inverses, products, axis rotations and a quotient:
    (F_i, F_j, S_ij) == (eps1, eps2, eps1eps2) up to FP32 evaluation order."""
    exe = str(tmp_path / "host_jet_hessian")
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-o", exe, os.path.join(ROOT, "tests", "harness", "host_jet_hessian.cpp")],
                   check=True)
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    assert len(out) == 9
    for line in out:
        i, j, e1, e2, e12 = line.split()
        assert float(e1) <= 1e-6 and float(e2) <= 1e-6 and float(e12) <= 1e-5, line
