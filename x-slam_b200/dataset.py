"""Depth-stream readers in front of the frame loop: the reference's Dataset API (XKinectFusion/include/Dataset.h:18-81,
src/Dataset.cpp:3-124) over the C-ABI of libxslam_b200.so (csrc/dataset.cpp: PNG decode without OpenCV).  Method names,
argument meaning and conventions follow the reference: frames `start_frame..end_frame` inclusive, `getDepthData`
returns uint16 millimetres (ICL raw / 5, optional horizontal flip), poses are 4x4 float32."""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import XsError


def _b(s):
    return s.encode() if isinstance(s, str) else s


def imread_depth(path):
    """cv::imread(path, cv::IMREAD_UNCHANGED) for the benchmarks' greyscale depth PNGs -> uint16 [rows, cols]."""
    lib = _capi.load()
    rows, cols = C.c_int(), C.c_int()
    _capi.check(lib.xs_read_png16(_b(path), None, 0, C.byref(rows), C.byref(cols)), "imread")
    out = np.empty((rows.value, cols.value), np.uint16)
    _capi.check(lib.xs_read_png16(_b(path), out.ctypes.data_as(C.POINTER(C.c_uint16)), out.size, C.byref(rows), C.byref(cols)), "imread")
    return out


def loadTxtMatrix(filename, rows, cols):
    """IOHelper.cpp:4-19."""
    out = np.zeros((rows, cols), np.float32)
    _capi.check(_capi.load().xs_load_txt_matrix(_b(filename), rows, cols, out.ctypes.data_as(C.POINTER(C.c_float))), "loadTxtMatrix")
    return out


class Dataset:
    """Dataset.h:18-61."""

    def __init__(self, handle=None):
        self._lib = _capi.load()
        self._h = handle

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.xs_dataset_close(self._h)
            self._h = None

    def size(self):
        return self._lib.xs_dataset_size(self._h) if self._h else 0

    def getDepthData(self, index, rows=480, cols=640):
        assert 0 <= index < self.size()
        out = np.empty((rows, cols), np.uint16)
        _capi.check(self._lib.xs_dataset_get_depth(self._h, index, out.ctypes.data_as(C.POINTER(C.c_uint16)), rows, cols), "getDepthData")
        return out

    def getPose(self, index):
        assert 0 <= index < self.size()
        out = np.zeros((4, 4), np.float32)
        _capi.check(self._lib.xs_dataset_get_pose(self._h, index, out.ctypes.data_as(C.POINTER(C.c_float))), "getPose")
        return out

    def getAllPose(self):
        return [self.getPose(i) for i in range(self.size())]

    def setPose(self, index, pose):
        assert 0 <= index < self.size()
        p = np.ascontiguousarray(pose, np.float32).reshape(16)
        _capi.check(self._lib.xs_dataset_set_pose(self._h, index, p.ctypes.data_as(C.POINTER(C.c_float))), "setPose")

    def getTimestamp(self, index):
        assert 0 <= index < self.size()
        return self._lib.xs_dataset_timestamp(self._h, index).decode()

    def depthFilename(self, index):
        return self._lib.xs_dataset_depth_filename(self._h, index).decode()


class ICL_Dataset(Dataset):
    """Dataset.h:75-81 / Dataset.cpp:69-124."""

    def __init__(self, dataset_dir, start_frame, end_frame, is_flip=False):
        super().__init__()
        self._h = self._lib.xs_dataset_open_icl(_b(dataset_dir), start_frame, end_frame, int(is_flip))
        if not self._h:
            raise XsError("ICL_Dataset: " + self._lib.xs_last_error().decode())

    @staticmethod
    def readPoseFile(poses_path, start, end):
        """Returns (ok, pose): lines [start, end) of the .gt.sim file -> top of a 4x4, last row 0 0 0 1."""
        out = np.zeros((4, 4), np.float32)
        rc = _capi.load().xs_icl_read_pose_file(_b(poses_path), start, end, out.ctypes.data_as(C.POINTER(C.c_float)))
        return rc == 0, out


class seven_scenes_Dataset(Dataset):
    """Dataset.h:63-73 / Dataset.cpp:13-67."""

    def __init__(self, dataset_dir, start_frames, end_frames, seq_names, is_flip=False):
        super().__init__()
        n = len(seq_names)
        s = (C.c_int * n)(*start_frames)
        e = (C.c_int * n)(*end_frames)
        names = (C.c_char_p * n)(*[_b(x) for x in seq_names])
        self._h = self._lib.xs_dataset_open_seven_scenes(_b(dataset_dir), s, e, names, n, int(is_flip))
        if not self._h:
            raise XsError("seven_scenes_Dataset: " + self._lib.xs_last_error().decode())

    @staticmethod
    def readInfo(filename, max_seq=64):
        """Returns (start_frames, end_frames, seq_names) - seq_names as "seq-XX/"."""
        s = (C.c_int * max_seq)()
        e = (C.c_int * max_seq)()
        names = C.create_string_buffer(16 * max_seq)
        n = _capi.load().xs_seven_scenes_read_info(_b(filename), s, e, names, max_seq)
        if n < 0:
            raise XsError("readInfo: " + _capi.load().xs_last_error().decode())
        raw = names.raw
        return list(s[:n]), list(e[:n]), [raw[16 * i:16 * (i + 1)].split(b"\0", 1)[0].decode() for i in range(n)]
