"""ctypes binding of libgcp.so (include/gcp.h).  No fallback: if the CUDA library is
missing or does not load, importing the engine raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# GCP_LIB=debug loads the bounds-checked build (build.py: libgcp_debug.so) - tests only;
# GCP_LIB=<tag> any other in-tree libgcp_<tag>.so (A/B runs of an experimental build)
_tag = os.environ.get("GCP_LIB", "")
LIB_PATH = os.path.join(HERE, "libgcp_%s.so" % _tag if _tag and _tag != "release" else "libgcp.so")

GCP_ABI_VERSION = 2
GCP_MAX_SHARDS = 16
GCP_MAX_AGENTS = 32
INVALID_EDGE = 0xFFFFFFFF


class GcpParams(C.Structure):
    _fields_ = [("num_agents", C.c_int32), ("max_dna_len", C.c_int32),
                ("solo_sep_thresh", C.c_int32), ("min_solo_lcs", C.c_int32),
                ("min_comb_lcs", C.c_int32), ("reserved0", C.c_int32),
                ("flight_time_scale", C.c_double), ("coverage_blend", C.c_double),
                ("one_minus_blend", C.c_double)]


class GcpVaryParams(C.Structure):
    _fields_ = [("min_dna_len", C.c_int32), ("crossover_time_thresh", C.c_int32),
                ("num_muterpolations", C.c_int32), ("muterpolation_srch_dist", C.c_int32),
                ("path_memory", C.c_int32), ("reserved0", C.c_int32),
                ("crossover_prob", C.c_double), ("mutation_prob", C.c_double),
                ("muterpolate_prob", C.c_double), ("muterpolation_sub_prob", C.c_double),
                ("log_scale_weight", C.c_double)]


class GcpMapConsts(C.Structure):
    _fields_ = [("height", C.c_int32), ("width", C.c_int32), ("nbx", C.c_int32),
                ("nby", C.c_int32), ("mask_size", C.c_int64), ("gsum_mask", C.c_int64),
                ("map_size", C.c_int64), ("num_occluded", C.c_double)]


class GcpFrustumParams(C.Structure):
    _fields_ = [("num_rays", C.c_int32), ("window", C.c_int32), ("scale", C.c_double),
                ("max_view", C.c_double), ("max_view_px", C.c_double), ("half_view", C.c_double),
                ("near_last", C.c_double * 2), ("near_first", C.c_double * 2),
                ("eps_angle", C.c_double), ("eps_round", C.c_double)]


class GcpGraphParams(C.Structure):
    _fields_ = [("scale", C.c_double), ("max_dist", C.c_double), ("num_dilations", C.c_int32),
                ("cell", C.c_int32), ("sectors", C.c_double * 8), ("eps_angle", C.c_double)]


class GcpShards(C.Structure):
    _fields_ = [("world", C.c_uint32), ("per_shard", C.c_uint32), ("base", C.c_void_p * 16)]


class GcpEvalOut(C.Structure):
    _fields_ = [("obj", C.c_void_p), ("neg_cov", C.c_void_p), ("dist", C.c_void_p),
                ("turn", C.c_void_p), ("cov", C.c_void_p), ("lc_mat", C.c_void_p),
                ("flags", C.c_void_p)]


# name -> (restype, argtypes); must list every symbol include/gcp.h declares
_vp, _u32, _u64, _i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
SIGNATURES = {
    "gcp_abi_version": (C.c_int, []),
    "gcp_create": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "gcp_destroy": (None, [_vp]),
    "gcp_last_error": (C.c_char_p, [_vp]),
    "gcp_upload_graph": (C.c_int, [_vp, _u32, _u32, _vp, _vp, _vp, _vp]),
    "gcp_upload_edges": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _u64, _vp, _u64]),
    "gcp_upload_masked_footprints": (C.c_int, [_vp, _vp, _vp, _vp, _u64, _vp, _u64]),
    "gcp_upload_map": (C.c_int, [_vp, C.POINTER(GcpMapConsts), _vp, _vp, _vp, _vp, _u64]),
    "gcp_set_params": (C.c_int, [_vp, C.POINTER(GcpParams)]),
    "gcp_eval_scratch_bytes": (C.c_size_t, [_vp, _u32]),
    "gcp_eval": (C.c_int, [_vp, _vp, _u32, C.POINTER(GcpEvalOut), _vp, C.c_size_t, _vp]),
    "gcp_path_costs": (C.c_int, [_vp, _vp, _u32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gcp_coverage_masked": (C.c_int, [_vp, _vp, _u32, _vp, _vp, _vp]),
    "gcp_coverage": (C.c_int, [_vp, _vp, _u32, _vp, _vp, _vp]),
    "gcp_loop_closures": (C.c_int, [_vp, _vp, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gcp_eval_host": (C.c_int, [_vp, _vp, _u32, _vp, _vp]),
    "gcp_eval_host16": (C.c_int, [_vp, _vp, _u32, _vp, _vp]),
    "gcp_eval16": (C.c_int, [_vp, _vp, _u32, C.POINTER(GcpEvalOut), _vp]),
    "gcp_eval16_available": (C.c_int, [_vp]),
    "gcp_host_narrow_active": (C.c_int, [_vp]),
    "gcp_rank_scratch_bytes": (C.c_size_t, [_vp, _u32]),
    "gcp_maximin": (C.c_int, [_vp, _vp, _u32, C.c_double, _u32, _u32, _vp, _vp, _vp,
                              C.c_size_t, _vp]),
    "gcp_arena_order": (C.c_int, [_vp, _vp, _u32, _vp, _vp, _vp, C.c_size_t, _vp]),
    "gcp_set_vary_params": (C.c_int, [_vp, C.POINTER(GcpVaryParams)]),
    "gcp_low_var_select": (C.c_int, [_vp, _vp, _u32, C.c_double, C.c_double, _vp, _vp, _vp]),
    "gcp_low_var_r0": (C.c_double, [_u64, _u32, _u32]),
    "gcp_crossover": (C.c_int, [_vp, _vp, _vp, _u32, _u64, _u32, _vp, _vp]),
    "gcp_mutate": (C.c_int, [_vp, _vp, _u32, _u64, _u32, _u32, _vp]),
    "gcp_select_vary": (C.c_int, [_vp, _vp, _vp, _u32, C.c_double, _u64, _u32, _vp, _vp,
                                  C.c_size_t, _vp]),
    "gcp_random_walks": (C.c_int, [_vp, _vp, _u32, _vp, C.c_int32, _u64, _u32, _vp]),
    "gcp_gather_survivors": (C.c_int, [_vp, _vp, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gcp_shard_alloc": (C.c_int, [_vp, C.c_size_t, C.POINTER(_vp), _vp]),
    "gcp_shard_free": (C.c_int, [_vp, _vp]),
    "gcp_shard_open": (C.c_int, [_vp, _vp, C.POINTER(_vp)]),
    "gcp_shard_close": (C.c_int, [_vp, _vp]),
    "gcp_crossover_shards": (C.c_int, [_vp, C.POINTER(GcpShards), C.c_int, _vp, _u32, _u32, _u32, _u64,
                                       _u32, _vp, _vp]),
    "gcp_mutate_genes": (C.c_int, [_vp, _vp, C.c_int, _u32, _u32, _u64, _u32, _u32, _vp]),
    "gcp_gather_shards": (C.c_int, [_vp, _vp, _u32, _u32, _u32, C.POINTER(GcpShards),
                                    C.POINTER(GcpShards), C.c_int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gcp_peer_barrier": (C.c_int, [_vp, C.POINTER(GcpShards), _u32, _u32, _vp, _vp]),
    "gcp_set_option": (C.c_int, [_vp, C.c_char_p, C.c_int64]),
    "gcp_launch_count": (C.c_uint64, [_vp]),
    "gcp_set_timing": (C.c_int, [_vp, C.c_int]),
    "gcp_stage_ms": (C.c_float, [_vp, C.c_int]),
    "gcp_build_graph": (C.c_int, [C.c_int, _vp, _u32, _u32, _vp, _u32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gcp_compute_frustums": (C.c_int, [C.c_int, _vp, _u32, _u32, _vp, _vp, _vp, _u32, _u32, _vp,
                                       _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
}

_lib = None


def load():
    """Load libgcp.so and bind every entry point.  Raises if the library is absent:
    build it with `python -m genetic_coverage_planning_b200.build`."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libgcp.so not built (%s); run `python -m "
                          "genetic_coverage_planning_b200.build` - there is no CPU fallback"
                          % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.gcp_abi_version() != GCP_ABI_VERSION:
        raise ImportError("libgcp.so ABI %d != binding %d" % (lib.gcp_abi_version(),
                                                              GCP_ABI_VERSION))
    _lib = lib
    return lib
