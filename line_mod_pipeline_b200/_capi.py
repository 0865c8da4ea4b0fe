"""ctypes binding of include/lmb200.h (the C ABI of liblmb200.so).  No torch types cross it."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("LMB200_SO") or os.path.join(_HERE, "liblmb200.so")  # LMB200_SO: A/B builds of the same library

MAX_MOD, MAX_LEVELS = 4, 8
T_8UC1, T_16UC1, T_8UC3 = 0, 2, 16
COLOR_GRADIENT, DEPTH_NORMAL = 0, 1
SIMLUT_CIRCULAR, SIMLUT_LINEAR = 0, 1
MB_L2_READ, MB_L1_READ, MB_HBM_READ, MB_H2D = range(4)
OK, E_INVALID, E_SOURCES, E_SIZE, E_FEATURES, E_CLASS, E_IO, E_CUDA, E_TRUNCATED, E_COMM, E_NODEVICE = range(0, -11, -1)
K_NAMES = ["upload", "pyrdown", "cg_quantize", "dn_quantize", "median", "decimate", "linearize",
           "sim_coarse", "sim_local", "pack", "comm", "epilogue"]
DBG_QUANTIZED, DBG_LINMEM, DBG_COARSE, DBG_UNSORTED, DBG_MAGNITUDE, DBG_DN_INDICES, DBG_SIMILARITY = range(7)


class Modality(C.Structure):
    _fields_ = [("type", C.c_int), ("weak_threshold", C.c_float), ("num_features", C.c_int),
                ("strong_threshold", C.c_float), ("distance_threshold", C.c_int),
                ("difference_threshold", C.c_int), ("extract_threshold", C.c_int)]


class Config(C.Structure):
    _fields_ = [("num_modalities", C.c_int), ("modalities", Modality * MAX_MOD), ("pyramid_levels", C.c_int),
                ("T", C.c_int * MAX_LEVELS), ("device", C.c_int), ("max_batch", C.c_int),
                ("candidate_capacity", C.c_int), ("similarity_lut", C.c_int)]


class Image(C.Structure):
    _fields_ = [("data", C.c_void_p), ("rows", C.c_int), ("cols", C.c_int), ("type", C.c_int), ("step", C.c_size_t)]


class Feature(C.Structure):
    _fields_ = [("x", C.c_int), ("y", C.c_int), ("label", C.c_int)]


class Template(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("pyramid_level", C.c_int), ("num_features", C.c_int),
                ("features", C.POINTER(Feature))]


class MatchRec(C.Structure):
    _fields_ = [("x", C.c_int), ("y", C.c_int), ("similarity", C.c_float), ("class_index", C.c_int),
                ("template_id", C.c_int)]


class Profile(C.Structure):
    _fields_ = [("ms", C.c_double * len(K_NAMES)), ("launches", C.c_longlong * len(K_NAMES)), ("bytes_coarse", C.c_longlong),
                ("bytes_local", C.c_longlong), ("frames", C.c_longlong), ("candidates", C.c_longlong),
                ("matches", C.c_longlong), ("chunks_coarse", C.c_longlong)]


class Mesh(C.Structure):
    _fields_ = [("vertices", C.POINTER(C.c_double)), ("n_vertices", C.c_int), ("triangles", C.POINTER(C.c_int)),
                ("n_triangles", C.c_int)]


class Camera(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double),
                ("cy", C.c_double), ("near_mm", C.c_double), ("far_mm", C.c_double)]


# every entry point include/lmb200.h declares: name -> (restype, argtypes)
_H = C.c_void_p
_P = C.POINTER
SIGNATURES = {
    "lmb200_render_lookat": (C.c_int, [_P(Mesh), _P(Camera), _P(C.c_double), C.c_int, _P(C.c_uint16), _P(C.c_uint8), C.c_int]),
    "lmb200_render_pose": (C.c_int, [_P(Mesh), _P(Camera), _P(C.c_double), _P(C.c_double), C.c_int, _P(C.c_uint16), _P(C.c_uint8), C.c_int]),
    "lmb200_hodan_error": (C.c_int, [_P(C.c_uint16), _P(C.c_uint16), _P(C.c_uint16), C.c_int, C.c_int, C.c_int, C.c_int, _P(C.c_float),
                                     _P(C.c_longlong), _P(C.c_longlong)]),
    "lmb200_hodan_error_poses": (C.c_int, [_P(Mesh), _P(Camera), _P(C.c_double), _P(C.c_double), _P(C.c_uint16), C.c_int, C.c_int, _P(C.c_float)]),
    "lmb200_load_ply": (C.c_int, [C.c_char_p, _P(_P(C.c_double)), _P(C.c_int), _P(_P(C.c_int)), _P(C.c_int)]),
    "lmb200_free": (None, [C.c_void_p]),
    "lmb200_default_modality": (None, [C.c_int, _P(Modality)]),
    "lmb200_default_config": (None, [_P(Config), C.c_int]),
    "lmb200_create": (C.c_int, [_P(Config), _P(_H)]),
    "lmb200_destroy": (None, [_H]),
    "lmb200_last_error": (C.c_char_p, [_H]),
    "lmb200_version": (C.c_char_p, []),
    "lmb200_num_modalities": (C.c_int, [_H]),
    "lmb200_modality_name": (C.c_char_p, [_H, C.c_int]),
    "lmb200_pyramid_levels": (C.c_int, [_H]),
    "lmb200_get_T": (C.c_int, [_H, C.c_int]),
    "lmb200_num_classes": (C.c_int, [_H]),
    "lmb200_class_id": (C.c_char_p, [_H, C.c_int]),
    "lmb200_num_templates": (C.c_int, [_H, C.c_char_p]),
    "lmb200_get_template": (C.c_int, [_H, C.c_char_p, C.c_int, C.c_int, _P(Template)]),
    "lmb200_add_template": (C.c_int, [_H, C.c_char_p, _P(Image), C.c_int, _P(Image), _P(C.c_int), _P(C.c_int)]),
    "lmb200_add_template_images": (C.c_int, [_H, C.c_char_p, _P(Image), C.c_int, _P(Image), _P(C.c_int), _P(C.c_int)]),
    "lmb200_add_template_pyramid": (C.c_int, [_H, C.c_char_p, _P(Template), C.c_int, _P(C.c_int)]),
    "lmb200_upload_templates": (C.c_int, [_H]),
    "lmb200_add_templates": (C.c_int, [_H, C.c_char_p, C.c_int, _P(Image), C.c_int, _P(Image), _P(C.c_int), _P(C.c_int)]),
    "lmb200_add_synthetic_template": (C.c_int, [_H, C.c_char_p, _P(Template), C.c_int, _P(C.c_int)]),
    "lmb200_clear_templates": (C.c_int, [_H]),
    "lmb200_write": (C.c_int, [_H, C.c_char_p]),
    "lmb200_read": (C.c_int, [C.c_char_p, C.c_int, _P(_H)]),
    "lmb200_write_class": (C.c_int, [_H, C.c_char_p, C.c_char_p]),
    "lmb200_read_class": (C.c_int, [_H, C.c_char_p, C.c_char_p]),
    "lmb200_write_classes": (C.c_int, [_H, C.c_char_p]),
    "lmb200_read_classes": (C.c_int, [_H, _P(C.c_char_p), C.c_int, C.c_char_p]),
    "lmb200_write_cache": (C.c_int, [_H, C.c_char_p]),
    "lmb200_read_cache": (C.c_int, [C.c_char_p, C.c_int, _P(_H)]),
    "lmb200_read_pose_sidecar": (C.c_int, [C.c_char_p, C.c_int, C.c_void_p, C.c_size_t, _P(C.c_size_t)]),
    "lmb200_write_pose_sidecar": (C.c_int, [C.c_char_p, _P(C.c_void_p), _P(C.c_size_t), C.c_int]),
    "lmb200_match": (C.c_int, [_H, _P(Image), C.c_int, C.c_float, _P(C.c_char_p), C.c_int, _P(MatchRec), C.c_size_t,
                               _P(C.c_size_t), _P(Image), _P(Image)]),
    "lmb200_match_batch": (C.c_int, [_H, _P(Image), C.c_int, C.c_int, C.c_float, _P(C.c_char_p), C.c_int,
                                     _P(MatchRec), C.c_size_t, _P(C.c_size_t)]),
    "lmb200_match_batch_submit": (C.c_int, [_H, _P(Image), C.c_int, C.c_int, C.c_float, _P(C.c_char_p), C.c_int, _P(C.c_int)]),
    "lmb200_match_batch_collect": (C.c_int, [_H, C.c_int, _P(MatchRec), C.c_size_t, _P(C.c_size_t)]),
    "lmb200_upload_frames": (C.c_int, [_H, _P(Image), C.c_int, C.c_int, C.c_int]),
    "lmb200_match_resident": (C.c_int, [_H, C.c_int, C.c_int, C.c_float, _P(C.c_char_p), C.c_int]),
    "lmb200_match_resident_sharded": (C.c_int, [_H, C.c_int, C.c_int, C.c_float, _P(C.c_char_p), C.c_int]),
    "lmb200_fetch_resident": (C.c_int, [_H, C.c_int, C.c_int, _P(MatchRec), C.c_size_t, _P(C.c_size_t)]),
    "lmb200_synchronize": (C.c_int, [_H]),
    "lmb200_stream": (C.c_void_p, [_H]),
    "lmb200_timer_record": (C.c_int, [_H, C.c_int]),
    "lmb200_timer_elapsed_ms": (C.c_int, [_H, _P(C.c_float)]),
    "lmb200_host_alloc": (C.c_int, [C.c_size_t, _P(C.c_void_p)]),
    "lmb200_host_free": (C.c_int, [C.c_void_p]),
    "lmb200_set_similarity_lut": (C.c_int, [_H, C.c_void_p]),
    "lmb200_get_similarity_lut": (C.c_int, [_H, C.c_void_p]),
    "lmb200_load_normal_lut": (C.c_int, [_H, C.c_char_p]),
    "lmb200_normal_lut_is_standin": (C.c_int, [_H]),
    "lmb200_warnings": (C.c_char_p, [_H]),
    "lmb200_set_normal_lut": (C.c_int, [_H, C.c_void_p]),
    "lmb200_get_normal_lut": (C.c_int, [_H, C.c_void_p]),
    "lmb200_set_template_shard": (C.c_int, [_H, C.c_int, C.c_int]),
    "lmb200_comm_unique_id": (C.c_int, [C.c_void_p]),
    "lmb200_comm_init": (C.c_int, [_H, C.c_void_p, C.c_int, C.c_int]),
    "lmb200_comm_destroy": (C.c_int, [_H]),
    "lmb200_fetch_resident_allgather": (C.c_int, [_H, C.c_int, C.c_int, _P(MatchRec), C.c_size_t, _P(C.c_size_t)]),
    "lmb200_merge_matches": (C.c_int, [_P(_P(MatchRec)), _P(C.c_size_t), C.c_int, _P(MatchRec), C.c_size_t, _P(C.c_size_t)]),
    "lmb200_debug_sort_check": (C.c_int, [_P(MatchRec), C.c_size_t, C.c_int, _P(MatchRec), _P(MatchRec)]),
    "lmb200_debug_shard_epilogue": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                              _P(MatchRec), C.c_size_t, C.c_void_p]),
    "lmb200_debug_merge_gathered": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                              _P(MatchRec), C.c_size_t, _P(C.c_size_t)]),
    "lmb200_shard_plan": (C.c_int, [_P(C.c_double), C.c_int, C.c_int, _P(C.c_int)]),
    "lmb200_postmatch_color": (C.c_int, [_H, C.c_int, C.c_void_p, C.c_void_p, _P(MatchRec), C.c_size_t, _P(C.c_int), _P(C.c_int)]),
    "lmb200_postmatch_median_depth": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_size_t, _P(C.c_int), C.c_int, _P(C.c_uint16)]),
    "lmb200_group_matches": (C.c_int, [_P(MatchRec), C.c_size_t, C.c_float, C.c_float, _P(C.c_int), _P(C.c_int)]),
    "lmb200_set_option": (C.c_int, [_H, C.c_char_p, C.c_int]),
    "lmb200_set_profiling": (C.c_int, [_H, C.c_int]),
    "lmb200_get_profile": (C.c_int, [_H, _P(Profile), C.c_int]),
    "lmb200_microbench": (C.c_int, [C.c_int, C.c_size_t, C.c_int, _P(C.c_double)]),
    "lmb200_debug_fetch": (C.c_int, [_H, C.c_int, C.c_int, C.c_int, C.c_void_p, _P(C.c_size_t)]),
}

_lib = None


def lib():
    """Loads liblmb200.so.  Fails loudly if it was not built: there is no Python/CPU fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError("liblmb200.so is missing — run `python -m line_mod_pipeline_b200.build` "
                              "(or __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(SO_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib
