"""B200-native (sm_100a, FP64) batched MSCKF measurement-update hot path.

Host side of the drop-in for the reference's `MSCKF.update` / `process_imu` path
(ValerioSpagnoli/Monocular-Visual-Inertial-MSCKF, src/msckf/MSCKF.py).  The compute lives in hand-written CUDA kernels
behind the C ABI of include/msckf_b200.h (csrc/libmsckf_b200.so, built in-tree by `make` / __graft_entry__.build()).
There is no CPU fallback: every compute entry point raises if the library or a CUDA device is missing.
"""
from ._lib import LibraryNotBuilt, lib_path, load_library  # noqa: F401
from .batched import BatchedMSCKF, FrameGraph, UpdateParams, pack_cases  # noqa: F401
from .dataset import MetricsAccumulator, project_landmarks, synth_imu  # noqa: F401
from .msckf import (B200UpdateMixin, Camera, Feature, IMU, IMUMeasurement, InverseDepthPoint, Isometry3D, Line,  # noqa: F401
                    MSCKF, MSCKFParameters, MSCKFState)

__all__ = ["BatchedMSCKF", "FrameGraph", "UpdateParams", "pack_cases", "synth_imu", "project_landmarks", "MetricsAccumulator", "load_library", "lib_path", "LibraryNotBuilt",
           "MSCKF", "MSCKFParameters", "MSCKFState", "B200UpdateMixin", "Feature", "Camera", "IMU", "IMUMeasurement",
           "Line", "InverseDepthPoint", "Isometry3D"]
