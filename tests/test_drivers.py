"""The C++ drivers (drivers/test_kinect_fusion.cpp, drivers/test_CSFD.cpp): the reference's Experiments/* surface.

CPU part: the binaries are built, link the library, read the reference's flat YAML and fail LOUDLY without a CUDA
device (no CPU fallback).  GPU part: outputs in the reference's formats (frame-%06d.pose.txt as IOHelper.cpp:21-32,
pcd.ply as CPointCloud.cpp:42-67), the test_CSFD known answers (main.cpp:194-219: 2.73911 / 9.26892), and agreement of
the driver's trajectory with the Python mirror of the same C-ABI."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "x-slam_b200", "bin")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_driver_binaries_exist_and_link():
    for name in ("test_kinect_fusion", "test_CSFD"):
        path = os.path.join(BIN, name)
        assert os.path.exists(path), "run `make -C x-slam_b200` (__graft_entry__.build())"
        out = subprocess.run(["ldd", path], capture_output=True, text=True).stdout
        assert "libxslam_b200.so" in out and "not found" not in out.split("libxslam_b200.so")[1].split("\n")[0]


def test_driver_usage_and_missing_key(tmp_path):
    r = subprocess.run([os.path.join(BIN, "test_kinect_fusion")], capture_output=True, text=True)
    assert r.returncode != 0 and "please enter the config file name" in r.stdout  # main.cpp:19-24
    bad = tmp_path / "bad.yaml"
    bad.write_text("dataset_format: synthetic\nstart_frame: 0\n")
    r = subprocess.run([os.path.join(BIN, "test_kinect_fusion"), str(bad)], capture_output=True, text=True)
    assert r.returncode != 0  # yaml-cpp throws on a missing key; so does the flat reader


@pytest.mark.skipif(_have_gpu(), reason="checks the no-GPU failure mode")
def test_drivers_fail_loudly_without_gpu(tmp_path):
    r = subprocess.run([os.path.join(BIN, "test_CSFD")], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr
    r = subprocess.run([os.path.join(BIN, "test_kinect_fusion"), os.path.join(ROOT, "configs", "synth_traj2.yaml"), str(tmp_path)],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


def test_flat_yaml_reads_reference_style_config(xs):
    cfg = xs.load_yaml(os.path.join(ROOT, "configs", "synth_traj2.yaml"))
    for key in ("tsdf_size_x", "tsdf_voxel_size", "max_integration_weight", "thres_range", "init_x", "r_x", "depth_width", "fx", "fy",
                "num_levels", "distThres", "angleThres", "biInterpolate_threshold", "start_frame", "end_frame", "log_slam_pose"):
        assert key in cfg, key  # the keys KinectFusionReconstruction.cpp:12-72 and main.cpp:28-33 read
    assert cfg["fy"] == -480.0 and cfg["log_slam_pose"] is True and cfg["dataset_format"] == "synthetic"


@pytest.mark.gpu
def test_csfd_driver_known_answers():
    r = subprocess.run([os.path.join(BIN, "test_CSFD")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    grads = [float(x) for x in re.findall(r"gradient = ([-0-9.eE+]+)", r.stdout)]
    seconds = [float(x) for x in re.findall(r"second order differentiation = ([-0-9.eE+]+)", r.stdout)]
    # a. DCSFD with the host scalar, b. chain rule from seeded partials, c. the same chain on the device arrays
    assert len(grads) == 3 and len(seconds) == 3
    for g_, s_ in zip(grads, seconds):
        assert abs(g_ - 2.73911) < 2e-3 and abs(s_ - 9.26892) < 0.1      # test_CSFD prints 2.73911 / 9.26892
    # host value pairs ("ours" beside std::complex) as main.cpp:113,132,151,171,191 prints them
    pairs = re.findall(r"(\w+) value: \(([-0-9.eE+]+),([-0-9.eE+]+)\)\t\(([-0-9.eE+]+),([-0-9.eE+]+)\)", r.stdout)
    want = {"multiplication": (-0.75, -1e-6), "division": (-0.333333, -8.88889e-7), "exp": (0.367879, 7.35759e-7),
            "sin": (-0.841471, 1.0806e-6), "pow": (-1.0, 6e-6)}
    assert [p_[0] for p_ in pairs] == list(want)
    for name, r0, i0, r1, i1 in pairs:
        wr, wi = want[name]
        assert abs(float(r1) - wr) < 2e-5 and abs(float(i1) - wi) < 2e-5 * abs(wi) + 1e-11, (name, r1, i1)
        assert abs(float(r0) - wr) < 2e-5 and abs(float(i0) - wi) < 0.06 * abs(wi), (name, r0, i0)  # "ours": pow prints 5.6982e-06
    # the device arrays give the same first-order values
    vals = re.findall(r"^value: \(([-0-9.eE+]+),([-0-9.eE+]+)\)", r.stdout, re.M)
    assert len(vals) == 5
    want["pow"] = (0.125, 0.75e-6)  # the device pow runs on a = (0.5, h)
    for (re_, im_), (wr, wi) in zip(vals, want.values()):
        assert abs(float(re_) - wr) < 2e-5 * max(1, abs(wr)) and abs(float(im_) - wi) < 2e-5 * max(abs(wi), 1e-7) + 1e-11, (re_, im_, wr, wi)


@pytest.mark.gpu
def test_kinect_fusion_driver_outputs(xs, tmp_path):
    cfg_text = open(os.path.join(ROOT, "configs", "synth_traj2.yaml")).read().replace("end_frame: 30", "end_frame: 4")
    cfg_path = tmp_path / "cfg.yaml"
    cfg_path.write_text(cfg_text)
    out = str(tmp_path / "out") + "/"
    r = subprocess.run([os.path.join(BIN, "test_kinect_fusion"), str(cfg_path), out], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr + r.stdout[-2000:]
    assert re.search(r"mean frame time = [0-9.]+ ms", r.stdout)  # main.cpp:83
    # trajectory files: 4 lines x 4 fixed-precision-7 values, each followed by a space (IOHelper.cpp:21-32)
    k = xs.KinectFusionReconstruction()
    k.SetYamlParameters(xs.load_yaml(str(cfg_path)), comps=1, seeds=xs.pose_seeds_csfd())
    for f in range(4):
        for sub in ("slam", "gt"):
            path = os.path.join(out, sub, "frame-%06d.pose.txt" % f)
            lines = open(path).read().split("\n")
            assert len(lines) == 5 and lines[4] == ""
            for ln in lines[:4]:
                assert re.fullmatch(r"(-?[0-9]+\.[0-9]{7} ){4}", ln), repr(ln)
        assert k.ProcessFrame(xs.synth_depth(f)) == 1
        slam = np.loadtxt(os.path.join(out, "slam", "frame-%06d.pose.txt" % f))
        assert np.abs(slam - k.pose_c2w()).max() < 5e-7  # same C-ABI, same frames: identical up to the printed precision
        gt = np.loadtxt(os.path.join(out, "gt", "frame-%06d.pose.txt" % f))
        assert np.abs(slam - gt).max() < 2e-2  # the tracker follows the synthetic trajectory
        d = np.loadtxt(os.path.join(out, "slam", "frame-%06d.dpose.txt" % f))
        assert d.shape == (6, 16)
        want = k.world2camera[1:].reshape(6, 16) / 1e-7
        assert np.abs(d - want).max() <= 1e-5 * max(1.0, np.abs(want).max())
    ply = open(os.path.join(out, "pcd.ply")).read().split("\n")
    assert ply[0] == "ply" and ply[1] == "format ascii 1.0" and ply[3].startswith("element vertex ")
    n = int(ply[3].split()[-1])
    assert n > 10000 and ply[10] == "end_header" and len(ply[11].split()) == 6


@pytest.mark.gpu
def test_kinect_fusion_driver_icl_dataset(xs, tmp_path):
    """dataset_format ICL (the reference driver's path, main.cpp:34-51): an ICL-NUIM-shaped tree written from the synthetic
    frames (raw = 5 x mm, `livingRoom1n.gt.sim`) must reproduce the synthetic run's trajectory bit for bit."""
    from common import write_png16_fast
    root = str(tmp_path / "icl") + "/"
    os.makedirs(root + "depth")
    lines = []
    for f in range(4):
        d = xs.synth_depth(f)
        assert int(d.max()) * 5 < 65536
        write_png16_fast(root + "depth/%d.png" % f, (d.astype(np.uint32) * 5).astype(np.uint16))
        P = xs.synth_pose(f)
        for r in range(3):
            lines.append(" ".join("%.9g" % v for v in P[r]))
        lines.append("")
    open(root + "livingRoom1n.gt.sim", "w").write("\n".join(lines) + "\n")
    base = open(os.path.join(ROOT, "configs", "synth_traj2.yaml")).read().replace("end_frame: 30", "end_frame: 3")
    outs = {}
    for fmt in ("synthetic", "ICL"):
        text = base.replace("dataset_format: synthetic", "dataset_format: " + fmt).replace('dataset_dir: ""', 'dataset_dir: "%s"' % root)
        if fmt == "synthetic":
            text = text.replace("end_frame: 3", "end_frame: 4")  # synthetic: frames [start, end); ICL: start..end inclusive
        cfg_path = tmp_path / (fmt + ".yaml")
        cfg_path.write_text(text)
        out = str(tmp_path / ("out_" + fmt)) + "/"
        r = subprocess.run([os.path.join(BIN, "test_kinect_fusion"), str(cfg_path), out], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr + r.stdout[-2000:]
        if fmt == "ICL":
            assert "pose path:" + root + "livingRoom1n.gt.sim" in r.stdout and "frame num: 4" in r.stdout
        outs[fmt] = out
    for f in range(3):  # the reference loop stops at frame_id == end_frame (main.cpp:45): 3 of the 4 ICL frames
        for sub in ("slam", "gt"):
            a = open(os.path.join(outs["ICL"], sub, "frame-%06d.pose.txt" % f)).read()
            b = open(os.path.join(outs["synthetic"], sub, "frame-%06d.pose.txt" % f)).read()
            if sub == "slam":
                assert a == b, (f, sub)
            else:
                assert np.abs(np.loadtxt(os.path.join(outs["ICL"], sub, "frame-%06d.pose.txt" % f)) -
                              np.loadtxt(os.path.join(outs["synthetic"], sub, "frame-%06d.pose.txt" % f))).max() < 2e-6
    assert not os.path.exists(os.path.join(outs["ICL"], "slam", "frame-000003.pose.txt"))
