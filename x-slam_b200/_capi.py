"""ctypes binding of libxslam_b200.so (include/xslam_b200.h).

The shared library is the product; this module only declares its C-ABI.  There is no fallback of any
kind: if the library is missing, `load()` raises, and every compute entry point returns XS_ERR_CUDA
when no CUDA device is present.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libxslam_b200.so")


class Intr(C.Structure):
    """Intr, Internal.h:49-59"""
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float)]

    def level(self, i):
        d = 1 << i
        return Intr(self.fx / d, self.fy / d, self.cx / d, self.cy / d)


class Pose(C.Structure):
    _fields_ = [("R", C.c_float * 9), ("t", C.c_float * 3), ("ncomp", C.c_int),
                ("dR", C.POINTER(C.c_float)), ("dt", C.POINTER(C.c_float))]


class Config(C.Structure):
    """The YAML keys of KinectFusionReconstruction::SetYamlParameters (KinectFusionReconstruction.cpp:12-72)."""
    _fields_ = [("res", C.c_int * 3), ("voxel_size", C.c_float), ("max_weight", C.c_int), ("thres_range", C.c_float),
                ("init_xyz", C.c_float * 3), ("r_deg", C.c_float * 3), ("width", C.c_int), ("height", C.c_int),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("num_levels", C.c_int),
                ("dist_thres", C.c_float), ("angle_thres_deg", C.c_float), ("bi_threshold", C.c_float),
                ("trunc_k", C.c_float), ("frame_step", C.c_int)]


# every symbol include/xslam_b200.h declares: name -> (restype, argtypes)
_vp, _i, _f, _l, _sz = C.c_void_p, C.c_int, C.c_float, C.c_long, C.c_size_t
_pf, _pd, _pi = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int)
_pull = C.POINTER(C.c_ulonglong)
_PP = C.POINTER(Pose)
SYMBOLS = {
    "xs_last_error": (C.c_char_p, []),
    "xs_version": (_i, []),
    "xs_launch_count": (C.c_longlong, []),
    "xs_dc_apply": (_i, [_i, _vp, _vp, _f, _vp, _l, _vp]),
    "xs_dc_chain": (_i, [_vp, _f, _vp, _l, _vp]),
    "xs_dc_host_apply": (_i, [_i, _vp, _vp, _f, _vp, _l]),
    "xs_dc_array_create": (_vp, [_l]),
    "xs_dc_array_release": (None, [_vp]),
    "xs_dc_array_resize": (_i, [_vp, _l]),
    "xs_dc_array_size": (_l, [_vp]),
    "xs_dc_array_ptr": (_vp, [_vp]),
    "xs_dc_array_upload": (_i, [_vp, _vp, _l]),
    "xs_dc_array_download": (_i, [_vp, _vp]),
    "xs_dc_array_copy": (_i, [_vp, _vp]),
    "xs_bilateral_filter": (_i, [_vp, _sz, _i, _i, _vp, _vp]),
    "xs_pyr_down": (_i, [_vp, _i, _i, _vp, _vp]),
    "xs_create_vmap": (_i, [Intr, _vp, _i, _i, _vp, _vp]),
    "xs_create_nmap": (_i, [_vp, _i, _i, _vp, _vp]),
    "xs_resize_vmap": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "xs_resize_nmap": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "xs_map_complex_to_soa": (_i, [_vp, _sz, _i, _i, _i, _vp, _i, _i, _vp]),
    "xs_map_soa_to_complex": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "xs_volume_create": (_vp, [_pi, _f, _f, _i, _i]),
    "xs_volume_create_hessian": (_vp, [_pi, _f, _f, _i, _i, _pi]),
    "xs_volume_last_raycast_hit_ms": (_f, [_vp]),
    "xs_volume_raycast_stats": (_i, [_vp, _pull]),
    "xs_volume_destroy": (None, [_vp]),
    "xs_volume_reset": (_i, [_vp, _vp]),
    "xs_volume_trunc_dist": (_f, [_vp]),
    "xs_volume_bytes": (_sz, [_vp]),
    "xs_volume_last_integrate_ms": (_f, [_vp]),
    "xs_volume_export_planes": (_i, [_vp, _i, _vp, _vp, _vp, _vp]),
    "xs_volume_import_planes": (_i, [_vp, _i, _vp, _vp, _vp, _vp]),
    "xs_integrate": (_i, [_vp, _vp, _sz, _i, _i, Intr, _i, _PP, _f, _pull, _vp]),
    "xs_volume_set_pipelined": (_i, [_vp, _i]),
    "xs_volume_finish_frame": (_i, [_vp, _pull]),
    "xs_raycast": (_i, [_vp, Intr, _PP, _PP, _i, _i, _vp, _vp, _vp]),
    "xs_tsdf_hessian": (_i, [_vp, _sz, _i, _i, Intr, _pi, _f, _PP, _f, _vp, _pd, _vp]),
    "xs_tsdf_hessian_batch": (_i, [_vp, _sz, _i, _i, Intr, _pi, _f, _PP, _f, _vp, _pd, _vp]),
    "xs_tsdf_loss": (_i, [_vp, _sz, _i, _i, Intr, _pi, _f, _pf, _pf, _f, _vp, _pd, _vp]),
    "xs_extract_points": (_l, [_vp, _vp, _vp, _l, _vp]),
    "xs_extract_normals": (_i, [_vp, _vp, _vp, _l, _vp]),
    "xs_estimate_combined": (_i, [_PP, _vp, _vp, _PP, Intr, _vp, _vp, _i, _i, _i, _i, _f, _f, _pd, _pd, _vp]),
    "xs_compute_optimize_matrix": (_l, [_PP, _vp, _vp, _PP, Intr, _vp, _vp, _i, _i, _f, _f, _pd, _pd, _vp]),
    "xs_kinfu_create": (_vp, [C.POINTER(Config), _i, _i, _pf, _i]),
    "xs_kinfu_create_hessian": (_vp, [C.POINTER(Config), _i, _i, _pi, _pf, _i]),
    "xs_kinfu_set_intrinsic_seeds": (_i, [_vp, _pf]),
    "xs_kinfu_keep_current_map_derivatives": (_i, [_vp, _i]),
    "xs_volume_set_intrinsic_seeds": (_i, [_vp, _pf]),
    "xs_kinfu_destroy": (None, [_vp]),
    "xs_kinfu_process_frame": (_i, [_vp, _vp, _i]),
    "xs_kinfu_set_deferred": (_i, [_vp, _i]),
    "xs_kinfu_sync": (_i, [_vp]),
    "xs_kinfu_surface_measure": (_i, [_vp, _vp]),
    "xs_kinfu_pose_estimate": (_i, [_vp]),
    "xs_kinfu_integrate_frame": (_i, [_vp, _vp]),
    "xs_kinfu_calculate_point_cloud": (_i, [_vp]),
    "xs_kinfu_frame_id": (_i, [_vp]),
    "xs_kinfu_get_world2camera": (_i, [_vp, _pf]),
    "xs_kinfu_get_pose_c2w": (_i, [_vp, _pf]),
    "xs_kinfu_volume": (_vp, [_vp]),
    "xs_kinfu_map": (_vp, [_vp, _i, _i, _pi, _pi, _pi]),
    "xs_kinfu_get_times": (_i, [_vp, _pf]),
    "xs_icp_deriv_times": (_i, [_vp, _vp, _i]),
    "xs_kinfu_enable_icp_log": (_i, [_vp, _i]),
    "xs_kinfu_take_icp_log": (_i, [_vp, _pd, _i]),
    "xs_kinfu_get_stats": (_i, [_vp, _pull]),
    "xs_kinfu_get_algorithmic_bytes": (_i, [_vp, _pd]),
    "xs_kinfu_pose_record_device": (_vp, [_vp]),
    "xs_kinfu_stream": (_vp, [_vp]),
    "xs_se3_exp": (_i, [_pf, _i, _i, _i, _pi, _pf]),
    "xs_kinfu_set_world2camera": (_i, [_vp, _pf]),
    "xs_kinfu_set_gt_poses": (_i, [_vp, _pf, _i, _i]),
    "xs_set_device": (_i, [_i]),
    "xs_comm_unique_id": (_i, [C.c_char_p]),
    "xs_comm_create": (_vp, [_i, _i, C.c_char_p]),
    "xs_comm_destroy": (None, [_vp]),
    "xs_comm_rank": (_i, [_vp]),
    "xs_comm_world": (_i, [_vp]),
    "xs_comm_all_gather": (_i, [_vp, _vp, _vp, _l, _vp]),
    "xs_kinfu_set_comm": (_i, [_vp, _vp, _i]),
    "xs_kinfu_get_gathered_records": (_i, [_vp, _pf]),
    "xs_kinfu_get_gathered_records_lagged": (_i, [_vp, _i, _pf]),
    "xs_kinfu_gathered_records_device": (_vp, [_vp]),
    "xs_save_pose_txt": (_i, [C.c_char_p, _pf]),
    "xs_export_ply": (_i, [C.c_char_p, _pf, _pf, _l]),
    "xs_synth_depth": (_i, [_pf, Intr, _i, _i, C.POINTER(C.c_uint16)]),
    "xs_synth_pose": (_i, [_i, _pf]),
    "xs_read_png16": (_i, [C.c_char_p, C.POINTER(C.c_uint16), _l, _pi, _pi]),
    "xs_load_txt_matrix": (_i, [C.c_char_p, _i, _i, _pf]),
    "xs_icl_read_pose_file": (_i, [C.c_char_p, _i, _i, _pf]),
    "xs_dataset_open_icl": (_vp, [C.c_char_p, _i, _i, _i]),
    "xs_dataset_open_seven_scenes": (_vp, [C.c_char_p, _pi, _pi, C.POINTER(C.c_char_p), _i, _i]),
    "xs_seven_scenes_read_info": (_i, [C.c_char_p, _pi, _pi, C.c_char_p, _i]),
    "xs_dataset_size": (_i, [_vp]),
    "xs_dataset_get_depth": (_i, [_vp, _i, C.POINTER(C.c_uint16), _i, _i]),
    "xs_dataset_get_pose": (_i, [_vp, _i, _pf]),
    "xs_dataset_set_pose": (_i, [_vp, _i, _pf]),
    "xs_dataset_timestamp": (C.c_char_p, [_vp, _i]),
    "xs_dataset_depth_filename": (C.c_char_p, [_vp, _i]),
    "xs_dataset_close": (None, [_vp]),
}

_lib = None


def load():
    """Loads libxslam_b200.so and types every exported symbol.  Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libxslam_b200.so is not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C x-slam_b200`.  There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class XsError(RuntimeError):
    pass


def check(rc, what=""):
    if rc is None or rc >= 0:
        return rc
    raise XsError("%s failed (%d): %s" % (what, rc, load().xs_last_error().decode()))
