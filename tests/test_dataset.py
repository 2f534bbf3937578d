"""Dataset readers: x-slam_b200/csrc/dataset.cpp through the C-ABI against the CPU restatement
(oracle/dataset_oracle.py) and - where the cv2 wheel is importable - against OpenCV itself, the un-vendored dependency
the reference decodes with (cv::imread / cv::Mat /= / cv::flip, Dataset.cpp:3-11).  Host code: runs without a GPU.
Integer work: everything is compared bit-exactly."""
import os
import struct
import zlib

import numpy as np
import pytest

from oracle import dataset_oracle as orc


def _paeth(a, b, c):
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    return a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)


def write_png16(path, img, filters=None, bit_depth=16, idat_split=1 << 30):
    """Minimal PNG encoder (test fixture writer): greyscale, chosen scan-line filter per row, IDAT split into chunks."""
    rows, cols = img.shape
    bpp = bit_depth // 8
    if bpp == 2:
        b = np.zeros((rows, cols * 2), np.uint8)
        b[:, 0::2], b[:, 1::2] = img >> 8, img & 255
    else:
        b = img.astype(np.uint8)
    raw = bytearray()
    prev = np.zeros(cols * bpp, np.int32)
    for y in range(rows):
        ft = (filters[y % len(filters)] if filters else 0)
        cur = b[y].astype(np.int32)
        line = np.zeros_like(cur)
        for i in range(cols * bpp):
            a = cur[i - bpp] if i >= bpp else 0
            up = prev[i]
            c = prev[i - bpp] if i >= bpp else 0
            pred = [0, a, up, (a + up) >> 1, _paeth(a, up, c)][ft]
            line[i] = (cur[i] - pred) & 255
        raw.append(ft)
        raw += bytes(line.astype(np.uint8))
        prev = cur

    def chunk(t, body):
        return struct.pack(">I", len(body)) + t + body + struct.pack(">I", zlib.crc32(t + body))
    z = zlib.compress(bytes(raw), 6)
    out = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", cols, rows, bit_depth, 0, 0, 0, 0))
    out += chunk(b"tEXt", b"Comment\0synthetic depth")
    for i in range(0, len(z), idat_split):
        out += chunk(b"IDAT", z[i:i + idat_split])
    out += chunk(b"IEND", b"")
    open(path, "wb").write(out)


def _depth_image(rng, rows, cols):
    y, x = np.mgrid[0:rows, 0:cols]
    img = (9000 + 4000 * np.sin(x / 17.0) * np.cos(y / 11.0) + rng.integers(0, 40, (rows, cols))).astype(np.uint16)
    img[rng.random((rows, cols)) < 0.02] = 0
    img[0, 0], img[0, 1], img[1, 0] = 65535, 0, 32768
    return img


@pytest.fixture(scope="module")
def ds():
    from xslam_b200 import dataset
    return dataset


@pytest.mark.parametrize("filters", [[0], [1], [2], [3], [4], [4, 1, 3, 2, 0]])
def test_png_decode_all_filters(ds, tmp_path, filters):
    rng = np.random.default_rng(len(filters) * 7 + filters[0])
    img = _depth_image(rng, 37, 53)
    p = str(tmp_path / "d.png")
    write_png16(p, img, filters, idat_split=777)
    assert np.array_equal(orc.png_decode_gray(p), img)
    assert np.array_equal(ds.imread_depth(p), img)


def test_png_decode_8bit_and_errors(ds, tmp_path, xs):
    img = (np.arange(24 * 31).reshape(24, 31) % 251).astype(np.uint16)
    p = str(tmp_path / "g8.png")
    write_png16(p, img, [4, 3], bit_depth=8)
    assert np.array_equal(ds.imread_depth(p), img)
    assert np.array_equal(orc.png_decode_gray(p), img)
    bad = str(tmp_path / "bad.png")
    data = bytearray(open(p, "rb").read())
    data[60] ^= 0x55  # corrupt a chunk body: the CRC check must reject it
    open(bad, "wb").write(bytes(data))
    with pytest.raises(xs.XsError):
        ds.imread_depth(bad)
    with pytest.raises(xs.XsError):
        ds.imread_depth(str(tmp_path / "missing.png"))


def test_png_decode_matches_opencv(ds, tmp_path):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    img = _depth_image(rng, 120, 160)
    p = str(tmp_path / "cv.png")
    assert cv2.imwrite(p, img)  # OpenCV's own encoder (adaptive filters)
    ref = cv2.imread(p, cv2.IMREAD_UNCHANGED)
    assert ref.dtype == np.uint16 and np.array_equal(ref, img)
    assert np.array_equal(ds.imread_depth(p), ref)
    assert np.array_equal(orc.png_decode_gray(p), ref)
    # the reference's `depth /= factor_` and flip (Dataset.cpp:8-10) through OpenCV's own arithmetic
    div = cv2.convertScaleAbs  # noqa: F841 (8-bit only; the 16-bit path below is what cv::Mat /= evaluates)
    cv_div = cv2.multiply(ref, 1.0 / 5)  # Mat * (1 / s): saturate_cast<ushort>(x * 0.2)
    assert np.array_equal(np.rint(ref.astype(np.float64) * (1.0 / 5)).astype(np.uint16), cv_div)
    assert np.array_equal(cv2.flip(cv_div, 1), cv_div[:, ::-1])


def _make_icl(root, rng, n, rows=48, cols=64):
    os.makedirs(os.path.join(root, "depth"))
    imgs, poses, lines = [], [], []
    for i in range(n):
        img = _depth_image(rng, rows, cols)
        write_png16(os.path.join(root, "depth", "%d.png" % i), img, [4, 2, 1])
        imgs.append(img)
        P = np.eye(4, dtype=np.float64)
        P[:3, :] = rng.standard_normal((3, 4))
        poses.append(P.astype(np.float32))
        for r in range(3):
            lines.append(" ".join("%.9g" % v for v in P[r]))
        lines.append("")
    open(os.path.join(root, "livingRoom1n.gt.sim"), "w").write("\n".join(lines) + "\n")
    return imgs, poses


@pytest.mark.parametrize("flip", [False, True])
def test_icl_dataset(ds, tmp_path, flip):
    rng = np.random.default_rng(11)
    root = str(tmp_path) + "/"
    imgs, poses = _make_icl(root, rng, 6)
    d = ds.ICL_Dataset(root, 1, 4, flip)  # frames 1..4 inclusive (Dataset.cpp:76)
    files, oposes, stamps = orc.icl_dataset(root, 1, 4)
    assert d.size() == 4 == len(files)
    for k in range(4):
        assert d.getTimestamp(k) == stamps[k] == str(1 + k)
        assert d.depthFilename(k) == files[k]
        want = orc.get_depth_data(files[k], 5, flip)
        got = d.getDepthData(k, 48, 64)
        assert got.dtype == np.uint16 and np.array_equal(got, want)
        raw5 = np.rint(imgs[1 + k].astype(np.float64) / 5).astype(np.uint16)
        assert np.array_equal(got, raw5[:, ::-1] if flip else raw5)
        assert np.array_equal(d.getPose(k), oposes[k])
        assert np.array_equal(d.getPose(k), poses[1 + k])
        assert np.array_equal(d.getPose(k)[3], [0, 0, 0, 1])
    ok, P = ds.ICL_Dataset.readPoseFile(root + "livingRoom1n.gt.sim", 8, 11)
    assert ok and np.array_equal(P, poses[2]) and np.array_equal(P, orc.icl_read_pose_file(root + "livingRoom1n.gt.sim", 8, 11))
    d.setPose(0, np.eye(4))
    assert np.array_equal(d.getPose(0), np.eye(4, dtype=np.float32))
    assert len(d.getAllPose()) == 4
    with pytest.raises(Exception):
        d.getDepthData(0, 480, 640)  # size mismatch is an error, not a silent reinterpretation


def test_icl_missing_pose_file(ds, tmp_path, xs):
    with pytest.raises(xs.XsError):
        ds.ICL_Dataset(str(tmp_path) + "/", 0, 1)


def test_seven_scenes_dataset(ds, tmp_path):
    rng = np.random.default_rng(5)
    root = str(tmp_path) + "/"
    imgs, poses = {}, {}
    for seq, (s, e) in (("01", (0, 2)), ("03", (998, 1000))):
        os.makedirs(os.path.join(root, "seq-" + seq))
        for f in range(s, e + 1):
            base = os.path.join(root, "seq-" + seq, "frame-%06d" % f)
            img = _depth_image(rng, 48, 64)
            write_png16(base + ".depth.png", img, [3, 4])
            P = rng.standard_normal((4, 4)).astype(np.float32)
            open(base + ".pose.txt", "w").write("\n".join("\t".join("%.7e" % v for v in row) + "\t " for row in P) + "\n")
            imgs[(seq, f)], poses[(seq, f)] = img, P
    open(root + "info.txt", "w").write("0 998\n2 1000\n01 03\n")
    s, e, names = ds.seven_scenes_Dataset.readInfo(root + "info.txt")
    assert (s, e, names) == orc.seven_scenes_read_info(root + "info.txt") == ([0, 998], [2, 1000], ["seq-01/", "seq-03/"])
    d = ds.seven_scenes_Dataset(root, s, e, names)
    files, oposes, stamps = orc.seven_scenes_dataset(root, s, e, names)
    assert d.size() == 6
    keys = [("01", 0), ("01", 1), ("01", 2), ("03", 998), ("03", 999), ("03", 1000)]
    for k, key in enumerate(keys):
        assert d.getTimestamp(k) == stamps[k] == "seq-%s/frame-%06d" % key
        assert np.array_equal(d.getDepthData(k, 48, 64), imgs[key])  # factor 1, no flip
        assert np.array_equal(d.getDepthData(k, 48, 64), orc.get_depth_data(files[k], 1, False))
        assert np.array_equal(d.getPose(k), oposes[k])
        np.testing.assert_allclose(d.getPose(k), poses[key], rtol=1e-6)
    assert np.array_equal(ds.loadTxtMatrix(files[0].replace(".depth.png", ".pose.txt"), 4, 4), oposes[0])
