"""line_mod_pipeline_b200 — B200-native (sm_100a CUDA) implementation of the cv::linemod::Detector::match
path that aelmiger/LINE-MOD-Pipeline drives from HighLevelLineMOD / PoseDetection.

The product is liblmb200.so (hand-written CUDA kernels behind the C ABI in include/lmb200.h); this
package holds its sources (csrc/), the build script and a thin ctypes mirror of the reference's
Detector interface.  There is no CPU fallback.
"""
from .detector import (Detector, ColorGradient, DepthNormal, getDefaultLINE, getDefaultLINEMOD, LinemodError,
                       MATCH_DTYPE, merge_matches, shard_plan, comm_unique_id, read_pose_sidecar, write_pose_sidecar, POSE_DTYPE, group_matches)
from . import _capi as capi
from . import render

__all__ = ["Detector", "ColorGradient", "DepthNormal", "getDefaultLINE", "getDefaultLINEMOD", "LinemodError",
           "MATCH_DTYPE", "merge_matches", "shard_plan", "comm_unique_id", "read_pose_sidecar", "write_pose_sidecar",
           "POSE_DTYPE", "group_matches", "capi", "render"]
