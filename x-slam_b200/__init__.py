"""xslam_b200 — B200-native CSFD/DCSFD-differentiated KinectFusion frame loop (host-side Python mirror).

The product is libxslam_b200.so (hand-written CUDA for sm_100a behind a C-ABI, include/xslam_b200.h).
This package mirrors the reference's operator and pipeline interface on top of it:

    ops.*                          the free functions of Map.h / TsdfFusion.h / RayCaster.h / ICP.h
    KinectFusionReconstruction     the pipeline class (SetYamlParameters, ProcessFrame, ...)
    dataset.*                      Dataset / ICL_Dataset / seven_scenes_Dataset (Dataset.h), PNG decode without OpenCV
    drivers.test_kinect_fusion     the YAML-driven demo driver (Experiments/test_xkinect_fusion/main.cpp)
    drivers.test_CSFD              the DCSFD self-check (Experiments/test_CSFD/main.cpp)

The directory is named `x-slam_b200`; import it as `xslam_b200` (repo-root shim xslam_b200.py).
"""
from . import _capi, dataset  # noqa: F401
from .dataset import Dataset, ICL_Dataset, seven_scenes_Dataset  # noqa: F401
from ._capi import Config, Intr, XsError, load  # noqa: F401
from .kinfu import (DEFAULT_CONFIG, H_, KinectFusionReconstruction, all_pairs, exportPly, hessian_seeds, load_yaml,  # noqa: F401
                    pose_seeds_csfd, pose_seeds_dcsfd, savePose, se3_exp, se3_generators, synth_depth, synth_pose)
