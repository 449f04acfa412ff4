"""ctypes binding of csrc/libmsckf_b200.so (C ABI: include/msckf_b200.h).  Fails loudly when the library is missing."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

COMPRESS_QR = 0
COMPRESS_GRAM = 1

FEAT_SKIPPED, FEAT_ACCEPTED, FEAT_GATED_OUT, FEAT_LOST_DISCARD, FEAT_LOST_FLAG = 0, 1, 2, 3, 0x10
FEAT_GAMMA_BOUND = 0x20
FEAT_TRIANGULATED = 0x40
ABI_VERSION = 10
GATE_EXACT, GATE_BOUNDS = 0, 1


class LibraryNotBuilt(RuntimeError):
    pass


class Dims(C.Structure):
    _fields_ = [("B", C.c_int), ("F", C.c_int), ("M", C.c_int), ("N", C.c_int), ("ldp", C.c_int)]


class Params(C.Structure):
    _fields_ = [("K", C.c_double * 9), ("Kinv", C.c_double * 9), ("width", C.c_int), ("height", C.c_int),
                ("sigma_image", C.c_double), ("gravity", C.c_double * 3), ("min_lost", C.c_int),
                ("min_tracked", C.c_int), ("use_parallax", C.c_int), ("min_parallax_deg", C.c_double),
                ("compress_mode", C.c_int), ("gate_mode", C.c_int)]


class DeviceBuffers(C.Structure):
    """msckf_device_buffers (include/msckf_b200.h): device pointers of the whole batch."""
    _fields_ = [(n, C.c_void_p) for n in (
        "P", "imu_state", "imu_null", "imu_time", "Qc_diag", "cam_R", "cam_t", "n_cams",
        "feat_len", "obs_uv", "obs_dir", "obs_conf", "obs_slot", "idp_base", "idp_anchor_slot", "idp_m", "idp_rho",
        "lost_cnt", "tracked_cnt", "feat_status", "gamma", "dof", "delta_x", "n_accepted", "n_rejected", "imu_samples")]


HOST_INPUTS = ("feat_len", "obs_uv", "obs_dir", "obs_conf", "obs_slot", "idp_base", "idp_anchor_slot", "idp_m",
               "idp_rho", "lost_cnt", "tracked_cnt")
HOST_OUTPUTS = ("imu_state", "cam_R", "cam_t", "idp_m_out", "idp_rho_out", "feat_status", "gamma", "delta_x",
                "n_accepted", "n_rejected")


class HostStep(C.Structure):
    """msckf_host_step (include/msckf_b200.h): host pointers of one camera-frame step."""
    _fields_ = ([(n, C.c_void_p) for n in HOST_INPUTS] +
                [("imu_samples", C.c_void_p), ("n_samples", C.c_int), ("use_null_state", C.c_int),
                 ("gravity", C.c_double * 3)] +
                [(n, C.c_void_p) for n in HOST_OUTPUTS])


TRACK_FIELDS = ("feat_len", "obs_uv", "obs_dir", "obs_conf", "obs_slot", "idp_base", "idp_anchor_slot", "idp_m", "idp_rho",
                "lost_cnt", "tracked_cnt")


class TrackArrays(C.Structure):   # msckf_track_arrays
    _fields_ = [(n, C.c_void_p) for n in TRACK_FIELDS]


# every symbol include/msckf_b200.h declares
EXPORTS = [
    "msckf_abi_version", "msckf_max_obs_per_feature", "msckf_max_cameras", "msckf_update_workspace_bytes",
    "msckf_update_batched", "msckf_front_batched", "msckf_compress_batched", "msckf_apply_update_batched",
    "msckf_propagate_batched", "msckf_augment_batched", "msckf_remove_workspace_bytes",
    "msckf_remove_cameras_batched", "msckf_launch_count", "msckf_probe_fp64", "msckf_filter_step_host",
    "msckf_set_stage_timing", "msckf_front_stage_ms", "msckf_select_poorest_cameras_batched",
    "msckf_drop_observations_batched", "msckf_append_observations_batched", "msckf_unobserved_cameras_batched",
    "msckf_compact_tracks_batched", "msckf_match_descriptors_batched", "msckf_verify_matches_batched",
    "msckf_synth_imu_batched", "msckf_project_landmarks_batched", "msckf_metrics_state_doubles",
    "msckf_metrics_step_batched", "msckf_metrics_finalize_batched",
    "msckf_score_filter_batched", "msckf_append_descriptors_batched", "msckf_drop_descriptors_batched",
    "msckf_compact_descriptors_batched", "msckf_track_descriptors_batched", "msckf_track_descriptors_from_sums_batched",
    "msckf_descriptor_rows_assign_batched", "msckf_descriptor_rows_remove_batched",
    "msckf_compact_tracks_inplace_batched", "msckf_compact_descriptors_inplace_batched",
]


def lib_path():
    """The in-tree library; MSCKF_B200_LIB names another build of the same ABI (A/B timing of kernel variants)."""
    return os.environ.get("MSCKF_B200_LIB") or os.path.join(_HERE, "csrc", "libmsckf_b200.so")


def load_library():
    """Load the CUDA library (once).  Raises LibraryNotBuilt -- there is no CPU fallback."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise LibraryNotBuilt(f"{path} is missing: run `make` (or __graft_entry__.build()) at the repo root. "
                              "This package has no CPU fallback.")
    lib = C.CDLL(path)
    vp, dp, ip, u8 = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p
    pd, pp = C.POINTER(Dims), C.POINTER(Params)
    lib.msckf_abi_version.restype = C.c_int
    lib.msckf_max_obs_per_feature.restype = C.c_int
    lib.msckf_max_cameras.restype = C.c_int
    lib.msckf_max_cameras.argtypes = [C.c_int]
    lib.msckf_launch_count.restype = C.c_longlong
    lib.msckf_update_workspace_bytes.restype = C.c_size_t
    lib.msckf_update_workspace_bytes.argtypes = [pd, C.c_int]
    lib.msckf_remove_workspace_bytes.restype = C.c_size_t
    lib.msckf_remove_workspace_bytes.argtypes = [pd]
    front_tail = [dp, dp, ip,                      # cam_R cam_t n_cams
                  ip, dp, dp, dp, ip, dp, ip, dp, dp, ip, ip,   # feat_len .. tracked_cnt
                  u8, dp, ip]                      # feat_status gamma dof
    lib.msckf_front_batched.restype = C.c_int
    lib.msckf_front_batched.argtypes = [pd, pp, dp, C.c_int, dp] + front_tail + [vp, C.c_size_t, vp]
    lib.msckf_update_batched.restype = C.c_int
    lib.msckf_update_batched.argtypes = ([pd, pp, dp, C.c_int, dp, dp] + front_tail +
                                         [dp, ip, ip, vp, C.c_size_t, vp])
    lib.msckf_compress_batched.restype = C.c_int
    lib.msckf_compress_batched.argtypes = [pd, pp, ip, ip, ip, u8, dp, dp, ip, ip, vp, C.c_size_t, vp]
    lib.msckf_apply_update_batched.restype = C.c_int
    lib.msckf_apply_update_batched.argtypes = [pd, pp, dp, dp, dp, dp, ip, dp, dp, ip, dp, vp, C.c_size_t, vp]
    lib.msckf_propagate_batched.restype = C.c_int
    lib.msckf_propagate_batched.argtypes = [pd, C.c_int, dp, dp, dp, C.POINTER(C.c_double), dp, dp, dp, ip, vp]
    lib.msckf_augment_batched.restype = C.c_int
    lib.msckf_augment_batched.argtypes = [pd, C.POINTER(C.c_double), dp, dp, dp, dp, ip, u8, ip, vp]
    lib.msckf_remove_cameras_batched.restype = C.c_int
    lib.msckf_remove_cameras_batched.argtypes = [pd, u8, dp, dp, dp, ip, vp, C.c_size_t, vp]
    lib.msckf_probe_fp64.restype = C.c_int
    lib.msckf_probe_fp64.argtypes = [C.c_int, C.c_int, dp, C.POINTER(C.c_double), vp]
    lib.msckf_filter_step_host.restype = C.c_int
    lib.msckf_filter_step_host.argtypes = [pd, pp, dp, C.c_int, C.POINTER(DeviceBuffers), C.POINTER(HostStep), C.c_int,
                                           C.c_int, vp, C.c_size_t, vp]
    lib.msckf_select_poorest_cameras_batched.restype = C.c_int
    lib.msckf_select_poorest_cameras_batched.argtypes = [pd, C.c_int, C.c_int, ip, ip, ip, u8, u8, vp]
    lib.msckf_unobserved_cameras_batched.restype = C.c_int
    lib.msckf_unobserved_cameras_batched.argtypes = [pd, ip, ip, ip, u8, u8, vp]
    lib.msckf_compact_tracks_batched.restype = C.c_int
    lib.msckf_compact_tracks_batched.argtypes = [pd, C.c_int, u8, C.POINTER(TrackArrays), C.POINTER(TrackArrays), u8, ip, vp]
    lib.msckf_compact_tracks_inplace_batched.restype = C.c_int
    lib.msckf_compact_tracks_inplace_batched.argtypes = [pd, C.c_int, u8, C.POINTER(TrackArrays), C.POINTER(TrackArrays), u8, ip, vp]
    lib.msckf_drop_observations_batched.restype = C.c_int
    lib.msckf_drop_observations_batched.argtypes = [pd, u8, ip, dp, ip, dp, dp, dp, ip, dp, ip, vp]
    lib.msckf_append_observations_batched.restype = C.c_int
    lib.msckf_append_observations_batched.argtypes = [pd, C.POINTER(C.c_double), ip, dp, dp, u8, dp, dp, ip, dp, dp, dp, ip, dp, ip,
                                                      dp, dp, ip, ip, ip, vp]
    lib.msckf_match_descriptors_batched.restype = C.c_int
    lib.msckf_match_descriptors_batched.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, ip, ip, vp, vp, ip, ip,
                                                    dp, dp, u8, dp, dp, ip, ip, vp]
    lib.msckf_verify_matches_batched.restype = C.c_int
    lib.msckf_verify_matches_batched.argtypes = [pd, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double, C.c_double,
                                                 ip, dp, dp, ip, dp, ip, dp, u8, u8, ip, ip, vp]
    lib.msckf_set_stage_timing.restype = C.c_int
    lib.msckf_set_stage_timing.argtypes = [C.c_int]
    lib.msckf_front_stage_ms.restype = C.c_int
    lib.msckf_front_stage_ms.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.msckf_synth_imu_batched.restype = C.c_int
    lib.msckf_synth_imu_batched.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, C.POINTER(C.c_double), dp, dp, C.c_ulonglong, C.c_int,
                                            C.c_int, dp, dp, dp, dp, vp]
    lib.msckf_project_landmarks_batched.restype = C.c_int
    lib.msckf_project_landmarks_batched.argtypes = [C.c_int, C.c_int, C.c_int, dp, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                                    C.c_int, C.c_int, dp, dp, C.c_double, C.c_ulonglong, dp, ip, ip, dp, vp]
    lib.msckf_metrics_state_doubles.restype = C.c_int
    lib.msckf_metrics_step_batched.restype = C.c_int
    lib.msckf_metrics_step_batched.argtypes = [pd, dp, dp, dp, dp, dp, vp]
    lib.msckf_metrics_finalize_batched.restype = C.c_int
    lib.msckf_metrics_finalize_batched.argtypes = [C.c_int, dp, dp, vp]
    lib.msckf_score_filter_batched.restype = C.c_int
    lib.msckf_score_filter_batched.argtypes = [C.c_int, C.c_int, C.c_int, ip, dp, dp, dp, dp, dp, dp, ip, vp]
    lib.msckf_append_descriptors_batched.restype = C.c_int
    lib.msckf_append_descriptors_batched.argtypes = [pd, C.c_int, C.c_int, u8, ip, ip, ip, ip, dp, ip, dp, dp, dp, C.c_int, ip, ip, vp]
    lib.msckf_drop_descriptors_batched.restype = C.c_int
    lib.msckf_drop_descriptors_batched.argtypes = [pd, C.c_int, u8, ip, ip, dp, dp, dp, C.c_int, ip, vp]
    lib.msckf_compact_descriptors_batched.restype = C.c_int
    lib.msckf_compact_descriptors_batched.argtypes = [pd, C.c_int, ip, ip, dp, dp, vp]
    lib.msckf_compact_descriptors_inplace_batched.restype = C.c_int
    lib.msckf_compact_descriptors_inplace_batched.argtypes = [pd, C.c_int, ip, ip, dp, dp, C.c_int, vp]
    lib.msckf_track_descriptors_batched.restype = C.c_int
    lib.msckf_track_descriptors_batched.argtypes = [pd, C.c_int, ip, dp, dp, u8, dp, dp, C.c_int, ip, ip, vp]
    lib.msckf_descriptor_rows_assign_batched.restype = C.c_int
    lib.msckf_descriptor_rows_assign_batched.argtypes = [pd, C.c_int, ip, u8, ip, vp]
    lib.msckf_descriptor_rows_remove_batched.restype = C.c_int
    lib.msckf_descriptor_rows_remove_batched.argtypes = [pd, ip, u8, ip, vp]
    lib.msckf_track_descriptors_from_sums_batched.restype = C.c_int
    lib.msckf_track_descriptors_from_sums_batched.argtypes = [pd, C.c_int, ip, dp, u8, dp, vp]
    if lib.msckf_abi_version() != ABI_VERSION:
        raise LibraryNotBuilt("libmsckf_b200.so ABI version mismatch; rebuild with `make`")
    _LIB = lib
    return lib


def check(rc, what):
    if rc != 0:
        kind = {-1: "invalid argument", -2: "capacity exceeded", -3: "workspace too small"}.get(rc, f"CUDA error {rc}")
        raise RuntimeError(f"{what} failed: {kind}")
