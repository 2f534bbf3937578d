"""configs[0] (test_CSFD): the CPU restatement of the reference's host bicomplex type is pinned against the known
answers the reference prints and against the reference's own DoubleComplex.cpp (oracle/_ref/libref_csfd.so)."""
import os
import subprocess

import numpy as np
import pytest

from common import rel_err


@pytest.fixture(scope="module")
def oracle():
    from oracle import pyref
    return pyref.Oracle()


def test_known_answers_of_test_CSFD(oracle):
    """Experiments/test_CSFD/main.cpp:203-219 at t = 0.5, h = 1e-6: gradient 2.73911, second order 9.26892."""
    h = 1e-6
    r = oracle.dc_chain(np.array([0.5]), h)[0]
    assert abs(r[1] / h - 2.73911) < 1e-4 and abs(r[2] / h - 2.73911) < 1e-4
    assert abs(r[3] / h / h - 9.26892) < 5e-2
    r64 = oracle.dc_chain(np.array([0.5]), h, f64=True)[0]
    t = 0.5
    f1 = 2 * (t * t + np.sin(t)) * (2 * t + np.cos(t))
    f2 = 2 * (2 * t + np.cos(t)) ** 2 + 2 * (t * t + np.sin(t)) * (2 - np.sin(t))
    assert abs(r64[1] / h - f1) < 1e-9 and abs(r64[3] / h / h - f2) < 1e-6


def test_value_prints_of_test_CSFD(oracle):
    """The five value pairs test_CSFD prints for a = (0.5, h), b = (-1.5, h)."""
    h = 1e-6
    a, b = complex(0.5, h), complex(-1.5, h)
    assert abs((a * b) - complex(-0.75, -1e-6)) < 1e-9
    assert abs((a / b) - complex(-0.333333, -8.88889e-07)) < 1e-6
    assert abs(np.exp(a + b) - complex(0.367879, 7.35759e-07)) < 1e-6
    assert abs(np.sin(a + b) - complex(-0.841471, 1.0806e-06)) < 1e-6


def test_against_reference_host_library(oracle):
    from oracle import pyref
    if not os.path.exists(pyref.REF_CSFD_PATH):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    ref = pyref.RefCsfd()
    rng = np.random.default_rng(0)
    n, h = 2000, 1e-6
    mk = lambda: np.stack([rng.uniform(0.5, 2.0, n), h * rng.standard_normal(n), h * rng.standard_normal(n),  # noqa: E731
                           h * h * rng.standard_normal(n)], 1).astype(np.float32)
    a, b = mk(), mk()
    for op in ("add", "sub", "mul", "div", "sqrt", "exp", "log", "sin", "cos", "pow"):
        bb = b if op in ("add", "sub", "mul", "div") else None
        r = ref.apply(op, a, bb, 3.0)
        m = oracle.dc_apply(op, a, bb, 3.0)
        assert np.array_equal(m, r) or rel_err(m, r) < 1e-6, op  # same std::complex<float> arithmetic
        m64 = oracle.dc_apply(op, a.astype(np.float64), None if bb is None else bb.astype(np.float64), 3.0, f64=True)
        for c in range(3):  # value and first-order parts of the FP32 reference against FP64
            assert rel_err(r[:, c], m64[:, c]) < 2e-5, (op, c)
    t = rng.uniform(0.1, 1.5, n).astype(np.float32)
    assert rel_err(oracle.dc_chain(t, h), ref.chain(t, h)) < 1e-6


def test_reference_binary_prints_known_answers():
    from oracle import pyref
    if not os.path.exists(pyref.REF_TEST_CSFD):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    out = subprocess.run([pyref.REF_TEST_CSFD], capture_output=True, text=True, timeout=120).stdout
    assert out.count("gradient = 2.73911") == 2 and out.count("second order differentiation = 9.26892") == 2
