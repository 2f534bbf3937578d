"""TEST INFRASTRUCTURE ONLY - CPU restatement of the reference's dataset readers, used by tests/ to
check x-slam_b200/csrc/dataset.cpp.  Follows XKinectFusion/src/Dataset.cpp:3-124 and src/IOHelper.cpp:4-19.

Third-party arithmetic: the reference decodes PNG and divides / flips through OpenCV (un-vendored, `find_package(OpenCV)`):
`cv::imread(path, IMREAD_UNCHANGED)`, `depth /= factor_` (cv::Mat arithmetic on CV_16U: x * (1 / factor), converted with
saturate_cast<ushort> = round to nearest, ties to even) and `cv::flip(depth, depth, 1)`.  The PNG decode here is the
published algorithm (PNG spec: zlib stream + five scan-line filters); it is pinned against OpenCV itself where the cv2
Python wheel is importable (tests/test_dataset.py) - the wheel bundles the same imgcodecs the C++ API uses."""
import os
import struct
import zlib

import numpy as np


def png_decode_gray(path):
    """Non-interlaced greyscale PNG (8 or 16 bit) -> uint16 [rows, cols]."""
    data = open(path, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n", "not a PNG"
    pos, idat, hdr = 8, b"", None
    while pos < len(data):
        n, typ = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        assert zlib.crc32(typ + body) == struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])[0], "CRC"
        if typ == b"IHDR":
            hdr = struct.unpack(">IIBBBBB", body)
        elif typ == b"IDAT":
            idat += body
        elif typ == b"IEND":
            break
        pos += 12 + n
    cols, rows, depth, color, _, _, interlace = hdr
    assert color == 0 and depth in (8, 16) and interlace == 0
    bpp = depth // 8
    stride = cols * bpp
    raw = zlib.decompress(idat)
    assert len(raw) == (stride + 1) * rows
    out = np.zeros((rows, stride), np.uint8)
    prev = np.zeros(stride, np.int32)
    for y in range(rows):
        ft = raw[(stride + 1) * y]
        line = np.frombuffer(raw, np.uint8, stride, (stride + 1) * y + 1).astype(np.int32)
        cur = np.zeros(stride, np.int32)
        if ft == 0:
            cur = line
        elif ft == 2:
            cur = (line + prev) & 255
        else:  # filters with a left neighbour are sequential
            for i in range(stride):
                a = cur[i - bpp] if i >= bpp else 0
                b = prev[i]
                c = prev[i - bpp] if i >= bpp else 0
                if ft == 1:
                    pred = a
                elif ft == 3:
                    pred = (a + b) >> 1
                else:
                    p = a + b - c
                    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
                    pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
                cur[i] = (line[i] + pred) & 255
        out[y] = cur
        prev = cur
    if bpp == 2:
        return (out[:, 0::2].astype(np.uint16) << 8) | out[:, 1::2].astype(np.uint16)
    return out.astype(np.uint16)


def get_depth_data(path, factor, is_flip):
    """Dataset::getDepthData, Dataset.cpp:3-11."""
    d = png_decode_gray(path)
    if factor != 1:
        d = np.rint(d.astype(np.float64) * (1.0 / factor)).astype(np.uint16)
    return d[:, ::-1].copy() if is_flip else d


def load_txt_matrix(path, rows, cols):
    """IOHelper.cpp:4-19."""
    vals = open(path).read().split()
    return np.array([float(v) for v in vals[:rows * cols]], np.float32).reshape(rows, cols)


def icl_read_pose_file(poses_path, start, end):
    """ICL_Dataset::readPoseFile, Dataset.cpp:90-124."""
    pose = np.eye(4, dtype=np.float32)
    for i, line in enumerate(open(poses_path).read().split("\n")):
        if start <= i < end:
            for j, sub in enumerate(line.split()):
                pose[i - start, j] = float(sub)
        elif i >= end:
            break
    pose[3] = [0, 0, 0, 1]
    return pose


def icl_dataset(dataset_dir, start_frame, end_frame):
    """ICL_Dataset::ICL_Dataset, Dataset.cpp:69-88 -> (depth filenames, poses, timestamps); factor = 5."""
    files, poses, stamps = [], [], []
    for i in range(start_frame, end_frame + 1):
        stamps.append(str(i))
        files.append(dataset_dir + "depth/" + str(i) + ".png")
        poses.append(icl_read_pose_file(dataset_dir + "livingRoom1n.gt.sim", 4 * i, 4 * i + 3))
    return files, poses, stamps


def seven_scenes_read_info(filename):
    """seven_scenes_Dataset::readInfo, Dataset.cpp:41-67."""
    lines = open(filename).read().split("\n")
    start = [int(x) for x in lines[0].split()]
    end = [int(x) for x in lines[1].split()]
    names = ["seq-" + x + "/" for x in lines[2].split()]
    return start, end, names


def seven_scenes_dataset(dataset_dir, start_frames, end_frames, seq_names):
    """seven_scenes_Dataset::seven_scenes_Dataset, Dataset.cpp:13-39; factor = 1."""
    files, poses, stamps = [], [], []
    for s, e, name in zip(start_frames, end_frames, seq_names):
        for frame in range(s, e + 1):
            base = name + "frame-" + "%06d" % frame
            stamps.append(base)
            files.append(dataset_dir + base + ".depth.png")
            poses.append(load_txt_matrix(dataset_dir + base + ".pose.txt", 4, 4))
    return files, poses, stamps
