"""ctypes binding of libnbw.so (include/nbw.h).  There is no CPU or PyTorch fallback: if the
CUDA extension has not been built, or no B200 is present, every compute entry raises."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnbw.so")
CSRC = os.path.join(_HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")
SOURCES = ["nbw_api.cu", "wl_relabel.cu", "dict_build.cu", "histogram.cu", "gram_i8.cu",
           "linalg_f64.cu", "gp_factor.cu", "score_pool.cu", "score_packed.cu", "topk.cu", "cellgen.cu", "pool_dedup.cu"]

NBW_MAX_LEVELS = 8
NBW_MAX_PARTS = 4
INVALID_ID = 0xFFFFFFFF

STATUS = {0: "OK", 1: "INVALID", 2: "CUDA", 3: "NOT_PD", 4: "HASH_COLLISION", 5: "CAPACITY",
          6: "UNSUPPORTED"}


class NbwError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libnbw: %s (%s)" % (msg, STATUS.get(code, code)))
        self.code = code


class NotPositiveDefinite(NbwError):
    pass


class HashCollision(NbwError):
    pass


class Model(C.Structure):
    """nbw_model (include/nbw.h)."""
    _fields_ = [
        ("n_max", C.c_int32), ("h", C.c_int32), ("oa", C.c_int32), ("acq", C.c_int32),
        ("seed", C.c_uint64),
        ("level0_table", C.c_void_p),
        ("table", C.c_void_p * NBW_MAX_LEVELS),
        ("table_mask", C.c_uint32 * NBW_MAX_LEVELS),
        ("col_off", C.c_int32 * NBW_MAX_LEVELS),
        ("label_off", C.c_int32 * NBW_MAX_LEVELS),
        ("thermo_base", C.c_void_p),
        ("max_count", C.c_void_p),
        ("n_features", C.c_int32), ("ld_g", C.c_int32),
        ("G", C.c_void_p), ("w_mu", C.c_void_p),
        ("n_labels", C.c_int32), ("ld_h", C.c_int32),
        ("H", C.c_void_p), ("w_mu_prefix", C.c_void_p),
        ("child", C.c_void_p), ("first_stable", C.c_int32), ("oa_d2_base", C.c_int32),
        ("lik", C.c_double), ("y_mean", C.c_double), ("y_std", C.c_double),
        ("incumbent", C.c_double), ("xi", C.c_double), ("beta", C.c_double),
        # packed scoring path (csrc/score_packed.cu)
        ("id_hash", C.c_void_p * NBW_MAX_LEVELS),
        ("bucket_tags", C.c_void_p * NBW_MAX_LEVELS),
        ("bucket_entries", C.c_void_p * NBW_MAX_LEVELS),
        ("bucket_mask", C.c_uint32 * NBW_MAX_LEVELS),
        ("final_label", C.c_void_p),
        ("level_weight", C.c_double * NBW_MAX_LEVELS),
        ("weighted", C.c_int32), ("undirected", C.c_int32), ("rec_offset", C.c_int32),
        ("rec_stride", C.c_int32), ("label_base", C.c_int32), ("compact", C.c_int32),
        ("palette", C.c_uint8 * 16),
        ("oa_exc", C.c_void_p),
        ("next", C.c_void_p),
    ]


_vp, _i64, _i32, _u64, _f64 = C.c_void_p, C.c_int64, C.c_int, C.c_uint64, C.c_double
_PROTOTYPES = {
    "nbw_version": (C.c_int, []),
    "nbw_create": (C.c_int, [C.POINTER(_vp), _i32]),
    "nbw_destroy": (C.c_int, [_vp]),
    "nbw_last_error": (C.c_char_p, [_vp]),
    "nbw_reserve": (C.c_int, [_vp, _u64]),
    "nbw_launch_count": (C.c_uint64, [_vp]),
    "nbw_profile_events": (C.c_int, [_vp, _vp, _vp]),
    "nbw_range_push": (C.c_int, [C.c_char_p]),
    "nbw_range_pop": (C.c_int, []),
    "nbw_wl_signatures": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _i32, _u64, _vp, _vp, _vp]),
    "nbw_dict_build": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _i64,
                                 C.POINTER(_i64), _vp]),
    "nbw_wl_fit": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _u64, _vp, _vp, _vp, C.POINTER(_i64), _vp]),
    "nbw_remap_ids": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "nbw_hist_dense_i8": (C.c_int, [_vp, _vp, _i64, _i32, _i32, C.POINTER(_i64), C.POINTER(_i64), _vp,
                                    _i32, _vp, _i64, _vp, _vp]),
    "nbw_oa_thermo_layout": (C.c_int, [_vp, _vp, _i64, _i32, _i32, C.POINTER(_i64), _vp, _vp,
                                       C.POINTER(_i64), _vp]),
    "nbw_gram_i8": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _i64, _i64, _i64, _i64, _vp, _i64, _i32, _vp]),
    "nbw_gram_i8_ex": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _i64, _i64, _i64, _i64, _vp, _i64, _i32, _i32, _i64, _vp, _vp,
                                 _vp, _vp]),
    "nbw_sym_mirror": (C.c_int, [_vp, _vp, _i64, _i64, _i32, _vp]),
    "nbw_gram_accumulate": (C.c_int, [_vp, _vp, _i64, _i64, _i64, _f64, _f64, _vp, _i64, _vp]),
    "nbw_axpby_f64": (C.c_int, [_vp, _vp, _i64, _i64, _i64, _f64, _f64, _vp, _i64, _vp]),
    "nbw_gram_normalize": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _vp]),
    "nbw_chol_f64": (C.c_int, [_vp, _vp, _i64, _i64, _f64, _vp, _i64, C.POINTER(_f64), C.POINTER(_i32), _vp]),
    "nbw_gp_factor": (C.c_int, [_vp, _vp, _i64, _i64, _f64, _vp, _vp, _i64, _vp, _i64, _vp,
                                C.POINTER(_f64), C.POINTER(_i32), _vp]),
    "nbw_gp_grid": (C.c_int, [_vp, _vp, _i64, _i64, _i64, C.POINTER(_i32), _i32, _f64, _vp, C.POINTER(_f64),
                              C.POINTER(_i32), _vp]),
    "nbw_trtri_lower_f64": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _i64, _vp]),
    "nbw_gemm_f64": (C.c_int, [_vp, _i32, _i32, _i64, _i64, _i64, _f64, _vp, _i64, _vp, _i64, _f64, _vp,
                               _i64, _vp]),
    "nbw_kernel_weight_grad": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _i64, _vp, _i64, C.POINTER(_vp), _i64, _i32,
                                         C.POINTER(C.c_double), _vp]),
    "nbw_scale_features": (C.c_int, [_vp, _vp, _i64, _i64, _i64, _i64, _vp, _f64, _vp, _i64, _vp]),
    "nbw_label_parents": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "nbw_level_grams_i32": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _i32, _vp, _i64, _i64, _vp]),
    "nbw_combine_columns": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _i64, _i32, _vp, _i64, _vp]),
    "nbw_prefix_features": (C.c_int, [_vp, _vp, _i64, _i64, _i32, C.POINTER(_i64), C.POINTER(_i64), _vp, _vp,
                                      _i64, _vp]),
    "nbw_posterior_cov_finish": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _i64, _f64, _f64, _vp, _i64, _vp]),
    "nbw_model_build_table": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _i64, _vp]),
    "nbw_model_build_id_hash": (C.c_int, [_vp, _vp, _i64, _u64, _i32, _vp, _vp]),
    "nbw_model_build_buckets": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp]),
    "nbw_score_pool": (C.c_int, [_vp, C.POINTER(Model), _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "nbw_score_pool_topk": (C.c_int, [_vp, C.POINTER(Model), _vp, _i64, _vp, _i32, _i32, _i64, _vp, _vp]),
    "nbw_score_pool_host": (C.c_int, [_vp, C.POINTER(Model), _vp, _i64, _vp, _vp, _i32, _vp]),
    "nbw_score_pool_mapped": (C.c_int, [_vp, C.POINTER(Model), _vp, _i64, _vp, _vp, _i32, _vp]),
    "nbw_wait_host_scores": (C.c_int, [_vp]),
    "nbw_pool_features": (C.c_int, [_vp, C.POINTER(Model), _vp, _i64, _vp, _i64, _vp, _vp]),
    "nbw_topk": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _i64, C.POINTER(C.c_float), C.POINTER(_i64),
                           C.POINTER(_i32), _vp]),
    "nbw_topk_merge_device": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "nbw_cands_read": (C.c_int, [_vp, _vp, _i32, C.POINTER(C.c_float), C.POINTER(_i64), C.POINTER(_i32), _vp]),
    "nbw_propose_host": (C.c_int, [_vp, C.POINTER(Model), _vp, _i64, _vp, _vp, _i32, _i32, _i32, _i32, _i64, _vp, _vp]),
    "nbw_topk_device": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _i64, _vp, _vp]),
    "nbw_topk_merge": (C.c_int, [_vp, _vp, _i64, _i32, _i32, C.POINTER(C.c_float), C.POINTER(_i64),
                                 C.POINTER(_i32), _vp]),
    "nbw_generate_cells": (C.c_int, [_vp, _i32, _u64, _i64, _i64, _vp, _vp]),
    "nbw_mutate_cells": (C.c_int, [_vp, _i32, _vp, _i64, _u64, _i64, _i64, _vp, _vp]),
    "nbw_pool_dedup": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _i64, _u64, _vp, C.POINTER(_i64), C.POINTER(_i64), _vp]),
    "nbw_generate_encoded": (C.c_int, [_vp, _i32, _u64, _i64, _i64, _vp, _vp, _vp]),
    "nbw_mutate_encoded": (C.c_int, [_vp, _i32, _vp, _i64, _u64, _i64, _i64, _i32, _vp, _vp, _vp]),
}
EXPORTED = tuple(_PROTOTYPES)

_lib = None


def nvcc_command(out_path: str = LIB_PATH):
    return ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
            "--threads", str(min(len(SOURCES), os.cpu_count() or 1)), "-shared", "-Xcompiler", "-fPIC", "-I" + INCLUDE, "-I" + CSRC, "-o", out_path] + \
           [os.path.join(CSRC, s) for s in SOURCES]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libnbw.so in-tree for sm_100a (cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(CSRC, "nbw_common.cuh"),
                                                       os.path.join(INCLUDE, "nbw.h")]
    if not force and os.path.exists(LIB_PATH) and \
            all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    cmd = nvcc_command()
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return LIB_PATH


def load():
    """Load libnbw.so; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "nasbowl_b200: %s is missing. Build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
