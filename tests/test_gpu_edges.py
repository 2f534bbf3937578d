"""Edge cases of the frame loop on the GPU against the reference's own kernels driven by the restated orchestrator
(oracle/_ref/libxslam_ref.so): empty and out-of-range depth, ragged image sizes (not multiples of the 32 x 8 tiles, odd
pyramid levels), a non-cubic volume, weight saturation, and loud failures for shapes the brick layout cannot hold.
Integer decisions (weights, validity masks, return codes) are compared bit-exactly; real parts to 1e-6."""
import numpy as np
import pytest

from common import H_, rel_err

pytestmark = pytest.mark.gpu


def _cfg(xs, w=640, h=480, res=(128, 128, 128), voxel=0.06, **kw):
    cfg = dict(xs.DEFAULT_CONFIG)
    s = w / 640.0
    cfg.update(tsdf_size_x=res[0], tsdf_size_y=res[1], tsdf_size_z=res[2], tsdf_voxel_size=voxel, depth_width=w, depth_height=h,
               fx=481.20 * s, fy=-480.00 * s, cx=(319.50 + 0.5) * s - 0.5, cy=(239.50 + 0.5) * h / 480.0 - 0.5)
    cfg.update(kw)
    return cfg


def _intr(cfg):
    return tuple(float(cfg[k]) for k in ("fx", "fy", "cx", "cy"))


def _compare_state(k, r, q=None):
    """weights / values / raycast masks of mine (k) vs a reference run (r); q: derivative component carried by r."""
    v, w, g = k.volume_planes(0)
    rv, rw, rg = r.volume()
    out = {"weight_mismatch": int((w.cpu().numpy() != rw).sum()), "value_rel": rel_err(v.cpu().numpy(), rv), "updated": int((rw > 0).sum())}
    for which in ("vmap_g_prev", "nmap_g_prev"):
        for level in range(3):
            m = k.map(which, level).cpu().numpy()
            rm = r.map(which, level)
            valid = ~np.isnan(rm[0, ..., 0])
            out["%s%d_mask" % (which[0], level)] = int((np.isnan(m[0, 0]) != ~valid).sum())
            both = valid & ~np.isnan(m[0, 0])
            out["%s%d_real" % (which[0], level)] = max(rel_err(m[0, p][both], rm[p, ..., 0][both]) for p in range(3)) if both.any() else 0.0
            out["%s%d_valid" % (which[0], level)] = int(valid.sum())
    return out


def test_empty_depth_frames(xs, refcuda):
    """All-zero depth: nothing is integrated, every raycast pixel is invalid, and the next frame's ICP has no
    correspondence - det(A) = 0 - so ProcessFrame returns 0 exactly like the reference (KinectFusionReconstruction.cpp:203-210)."""
    cfg = _cfg(xs)
    k = xs.KinectFusionReconstruction()
    k.SetYamlParameters(cfg, comps=1, seeds=xs.pose_seeds_csfd()[:2])
    r = refcuda.kinfu(cfg, None)
    empty = np.zeros((480, 640), np.uint16)
    assert k.ProcessFrame(empty) == 1 and r.process_frame(empty) == 1  # frame 0 has no ICP
    st = _compare_state(k, r)
    assert st["updated"] == 0 and st["weight_mismatch"] == 0 and st["v0_valid"] == 0 and st["v0_mask"] == 0 and st["n2_mask"] == 0
    assert r.process_frame(xs.synth_depth(1)) == 0
    assert k.ProcessFrame(xs.synth_depth(1)) == 0
    assert k.frame_id == 1  # the failed frame is not counted (ProcessFrame returns before frame_id += frame_step)


def test_out_of_range_depth_is_masked(xs, refcuda):
    """Depth outside the validity gates (0, < 200 mm, > 5000 mm; Map.cu:193, TsdfFusion.cu:77) in large patches and single
    pixels: weights, values and every validity mask of the pyramid match the reference bit for bit over two frames."""
    cfg = _cfg(xs)
    k = xs.KinectFusionReconstruction()
    k.SetYamlParameters(cfg, comps=1, seeds=xs.pose_seeds_csfd()[:1])
    r = refcuda.kinfu(cfg, None)
    rng = np.random.default_rng(7)
    for f in range(2):
        d = xs.synth_depth(f).copy()
        d[100:180, 200:330] = 0
        d[300:340, 50:120] = 150      # below the near gate
        d[20:60, 500:600] = 6000      # beyond the far gate
        d[400:470, 400:401] = 65535
        d[rng.random(d.shape) < 0.01] = 0
        assert k.ProcessFrame(d) == 1 and r.process_frame(d) == 1
    st = _compare_state(k, r)
    assert st["updated"] > 30000 and st["weight_mismatch"] == 0 and st["value_rel"] <= 1e-6, st
    assert all(st["%s%d_mask" % (m, l)] == 0 for m in "vn" for l in range(3)), st
    assert all(st["%s%d_real" % (m, l)] <= 1e-6 for m in "vn" for l in range(3)), st
    assert np.abs(k.world2camera[0] - r.pose().real).max() <= 1e-5


@pytest.mark.parametrize("w,h", [(168, 124), (100, 76)])
def test_ragged_image_sizes(xs, refcuda, w, h):
    """Image sizes that are not multiples of the 32 x 8 thread tiles (168 x 124 -> 84 x 62 -> 42 x 31; 100 x 76 -> 50 x 38 -> 25 x 19):
    two frames with one seeded direction, mine vs the reference kernels."""
    cfg = _cfg(xs, w, h, res=(64, 64, 64), voxel=0.12)
    seed = xs.pose_seeds_csfd()[4:5]
    k = xs.KinectFusionReconstruction()
    k.SetYamlParameters(cfg, comps=1, seeds=seed)
    r0 = refcuda.kinfu(cfg, None)
    r1 = refcuda.kinfu(cfg, seed[0].reshape(4, 4))
    for f in range(2):
        d = xs.synth_depth(f, w, h, *_intr(cfg))
        assert d.shape == (h, w)
        assert k.ProcessFrame(d) == 1 and r0.process_frame(d) == 1 and r1.process_frame(d) == 1
        st = _compare_state(k, r0)
        if f == 0:  # identical inputs to every stage on the first frame: bit-exact masks and weights
            assert st["weight_mismatch"] == 0 and st["value_rel"] <= 1e-6, st
            assert all(st["%s%d_mask" % (m, l)] == 0 for m in "vn" for l in range(3)), st
            assert st["v0_valid"] > 0.5 * w * h and st["v2_valid"] > 0
    assert np.abs(k.world2camera[0] - r0.pose().real).max() <= 1e-5
    sc = np.abs(r1.pose().imag).max()
    assert np.abs(k.world2camera[1] - r1.pose().imag).max() <= 2e-2 * sc
    assert st["weight_mismatch"] <= 1e-3 * st["updated"], st


def test_non_cubic_volume_and_weight_saturation(xs, refcuda):
    """A 128 x 64 x 96 volume and max_integration_weight = 2 over four frames: the weight plane saturates at 2 exactly where the
    reference's does (min(w + 1, max_weight), TsdfFusion.cu:167)."""
    cfg = _cfg(xs, res=(128, 64, 96), voxel=0.06, max_integration_weight=2, init_x=3.6, init_y=1.9, init_z=2.6)
    k = xs.KinectFusionReconstruction()
    k.SetYamlParameters(cfg, comps=3, seeds=xs.pose_seeds_dcsfd([(0, 4)])[0])
    r = refcuda.kinfu(cfg, None)
    for f in range(4):
        d = xs.synth_depth(f)
        assert k.ProcessFrame(d) == 1 and r.process_frame(d) == 1
    _, w, _ = k.volume_planes(0)
    _, rw, _ = r.volume()
    w = w.cpu().numpy()
    assert w.shape == (96, 64, 128) == rw.shape
    assert int(rw.max()) == 2 and int((rw == 2).sum()) > 10000
    assert int((w != rw).sum()) <= 1e-3 * int((rw > 0).sum())
    assert np.abs(k.world2camera[0] - r.pose().real).max() <= 2e-5


def test_unsupported_shapes_fail_loudly(xs):
    """The brick layout needs resolutions that are multiples of 8 and the pyramid needs even image sizes at every level it
    halves: anything else is an error at creation, never a silent truncation."""
    k = xs.KinectFusionReconstruction()
    with pytest.raises(xs.XsError):
        k.SetYamlParameters(_cfg(xs, res=(100, 128, 128)))
    with pytest.raises(ValueError):
        k.SetYamlParameters(_cfg(xs))
        k.ProcessFrame(np.zeros((479, 640), np.uint16))


def test_maximum_volume_1024(xs, refcuda):
    """BASELINE.json configs[4] resolution: a 1024^3 volume (0.0075 m voxels, 4 GiB per plane, derivative-plane offsets beyond
    2^31 elements) with one first-order direction.  One fused frame: the weight plane and the raycast validity masks must equal
    the reference kernels' bit for bit, the values to 1e-6."""
    import torch
    cfg = _cfg(xs, res=(1024, 1024, 1024), voxel=0.0075)
    seed = xs.pose_seeds_csfd()[1:2]
    k = xs.KinectFusionReconstruction()
    k.SetYamlParameters(cfg, comps=1, seeds=seed)
    r = refcuda.kinfu(cfg, None)
    d = xs.synth_depth(0)
    assert k.ProcessFrame(d) == 1 and r.process_frame(d) == 1
    v, w, g = k.volume_planes(0)
    rv, rw, _ = r.volume()
    rw_t = torch.from_numpy(rw).cuda()
    assert int((rw_t > 0).sum()) > 10_000_000
    assert bool(torch.equal(w, rw_t)), "updated-voxel sets differ at 1024^3"
    del rw_t
    rv_t = torch.from_numpy(rv).cuda()
    assert float((v - rv_t).abs().max()) <= 1e-6
    assert float(g.abs().max()) > 0  # the seeded direction reached the derivative plane
    del rv_t, v, w, g
    for which in ("vmap_g_prev", "nmap_g_prev"):
        m = k.map(which, 0).cpu().numpy()
        rm = r.map(which, 0)
        valid = ~np.isnan(rm[0, ..., 0])
        assert valid.sum() > 200000 and np.array_equal(np.isnan(m[0, 0]), ~valid)
        assert max(rel_err(m[0, p][valid], rm[p, ..., 0][valid]) for p in range(3)) <= 1e-6
