"""Shared helpers of the parity tests (inputs, pose construction, comparison metrics)."""
import numpy as np

ICL = dict(fx=481.20, fy=-480.00, cx=319.50, cy=239.50)
H_ = 1e-7


def world2volume(init=(3.2, 3.2, 3.2)):
    T = np.eye(4)
    T[:3, 3] = init
    return T


def rand_dpose(rng, n, scale=H_):
    """n h-scaled derivative components of a rigid transform: dR [n,9], dt [n,3] (generic, not necessarily
    tangent to SE(3): both implementations treat them as plain complex numbers)."""
    return (scale * rng.standard_normal((n, 9))).astype(np.float32), (scale * rng.standard_normal((n, 3))).astype(np.float32)


def poses_for_frame(xs, frame, init=(3.2, 3.2, 3.2)):
    """float32 real poses used by integration / raycast for a synthetic frame: (v2c, c2v, v2w) as 4x4 float64."""
    c2w = xs.synth_pose(frame).astype(np.float64)
    w2v = world2volume(init)
    c2v = w2v @ c2w
    return np.linalg.inv(c2v), c2v, np.linalg.inv(w2v)


def rel_err(a, b, floor=0.0):
    """max |a-b| / max(|b|) over finite entries (scale-relative error)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    m = np.isfinite(a) & np.isfinite(b)
    if not m.any():
        return 0.0
    scale = max(np.abs(b[m]).max(), floor, 1e-300)
    return float(np.abs(a[m] - b[m]).max() / scale)


def ulp_diff(a, b):
    """element-wise distance in float32 ulps (finite entries only)."""
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    ia = a.view(np.int32).astype(np.int64)
    ib = b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7fffffff), ia)
    ib = np.where(ib < 0, -(ib & 0x7fffffff), ib)
    d = np.abs(ia - ib)
    return np.where(np.isfinite(a) & np.isfinite(b), d, 0)


def write_png16_fast(path, img):
    """16-bit greyscale PNG with the Up filter on every row (vectorised; fixture writer for full-size frames)."""
    import struct
    import zlib
    rows, cols = img.shape
    b = np.zeros((rows, cols * 2), np.uint8)
    b[:, 0::2], b[:, 1::2] = img >> 8, img & 255
    prev = np.vstack([np.zeros((1, cols * 2), np.uint8), b[:-1]])
    lines = (b.astype(np.int16) - prev.astype(np.int16)).astype(np.uint8)
    raw = np.hstack([np.full((rows, 1), 2, np.uint8), lines]).tobytes()

    def chunk(t, body):
        return struct.pack(">I", len(body)) + t + body + struct.pack(">I", zlib.crc32(t + body))
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", cols, rows, 16, 0, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(raw, 1)) + chunk(b"IEND", b""))
