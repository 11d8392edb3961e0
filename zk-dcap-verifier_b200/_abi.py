"""ctypes image of include/zkdcap_b200.h.  No torch types cross this boundary."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ZKDCAP_B200_LIB") or os.path.join(_HERE, "libzkdcap_b200.so")   # override: A/B builds of the kernels


class ZkError(RuntimeError):
    pass


class zk_fr(C.Structure):
    _fields_ = [("l", C.c_uint64 * 4)]


class zk_value_source(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("a", C.c_uint32), ("b", C.c_uint32)]


class zk_calculation(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("a", zk_value_source), ("b", zk_value_source),
                ("parts_off", C.c_uint32), ("parts_len", C.c_uint32), ("target", C.c_uint32)]


class zk_graph(C.Structure):
    _fields_ = [("constants", C.c_void_p), ("n_constants", C.c_uint32),
                ("rotations", C.POINTER(C.c_int32)), ("n_rotations", C.c_uint32),
                ("calcs", C.POINTER(zk_calculation)), ("n_calcs", C.c_uint32),
                ("horner_parts", C.POINTER(zk_value_source)), ("n_horner_parts", C.c_uint32),
                ("num_intermediates", C.c_uint32)]


class zk_column(C.Structure):
    _fields_ = [("column_type", C.c_uint32), ("index", C.c_uint32)]


class zk_evaluator_desc(C.Structure):
    _fields_ = [("k", C.c_uint32), ("degree", C.c_uint32), ("blinding_factors", C.c_uint32),
                ("n_fixed", C.c_uint32), ("n_advice", C.c_uint32), ("n_instance", C.c_uint32),
                ("n_challenges", C.c_uint32),
                ("custom_gates", zk_graph),
                ("lookups", C.POINTER(zk_graph)), ("n_lookups", C.c_uint32),
                ("perm_columns", C.POINTER(zk_column)), ("n_perm_columns", C.c_uint32)]


class zk_query(C.Structure):
    _fields_ = [("column", C.c_uint32), ("rotation", C.c_int32)]


class zk_circuit_desc(C.Structure):
    _fields_ = [("ev", zk_evaluator_desc),
                ("lookup_inputs", C.POINTER(zk_graph)), ("lookup_tables", C.POINTER(zk_graph)),
                ("advice_queries", C.POINTER(zk_query)), ("n_advice_queries", C.c_uint32),
                ("fixed_queries", C.POINTER(zk_query)), ("n_fixed_queries", C.c_uint32),
                ("transcript_repr", zk_fr),
                ("instance_queries", C.POINTER(zk_query)), ("n_instance_queries", C.c_uint32)]


class zk_verifier_params(C.Structure):
    _fields_ = [("fixed_commitments", C.c_void_p), ("permutation_commitments", C.c_void_p),
                ("g1", C.c_uint64 * 8), ("g2", C.c_void_p), ("s_g2", C.c_void_p), ("transcript_kind", C.c_int)]


class zk_kernel_stat(C.Structure):
    _fields_ = [("name", C.c_char * 40), ("launches", C.c_uint64), ("total_ms", C.c_double)]


# ValueSource / Calculation / column kinds
VS_CONSTANT, VS_INTERMEDIATE, VS_FIXED, VS_ADVICE, VS_INSTANCE, VS_CHALLENGE, VS_BETA, VS_GAMMA, VS_THETA, VS_Y, \
    VS_PREVIOUS_VALUE = range(11)
CALC_ADD, CALC_SUB, CALC_MUL, CALC_SQUARE, CALC_DOUBLE, CALC_NEGATE, CALC_HORNER, CALC_STORE = range(8)
COL_ADVICE, COL_FIXED, COL_INSTANCE = range(3)
BASES_HOST, BASES_G, BASES_G_LAGRANGE = range(3)
TRANSCRIPTS = {"blake2b": 0, "evm": 1, "poseidon": 2}

# every symbol include/zkdcap_b200.h declares
SYMBOLS = [
    "zk_ctx_create", "zk_ctx_destroy", "zk_last_error", "zk_ctx_sync", "zk_version",
    "zk_ctx_timing", "zk_ctx_timing_reset", "zk_ctx_timing_read", "zk_ctx_launch_count",
    "zk_dev_alloc", "zk_dev_free", "zk_h2d", "zk_d2h",
    "zk_srs_upload", "zk_srs_upload_dev", "zk_g1_fixed_base_mul_dev", "zk_ctx_share_srs", "zk_srs_window", "zk_ctx_counters",
    "zk_params_read", "zk_params_write", "zk_g1_compress", "zk_g1_decompress",
    "zk_g1_encode", "zk_g1_decode", "zk_params_read_format", "zk_params_write_format",
    "zk_msm", "zk_msm_dev", "zk_msm_batch_dev", "zk_g1_sum", "zk_g1_sum_dev",
    "zk_ntt", "zk_ntt_dev",
    "zk_domain_create", "zk_domain_destroy", "zk_domain_k", "zk_domain_extended_k", "zk_domain_constants",
    "zk_lagrange_to_coeff", "zk_coeff_to_lagrange", "zk_coeff_to_extended", "zk_extended_to_coeff",
    "zk_divide_by_vanishing_poly",
    "zk_lagrange_to_coeff_dev", "zk_coeff_to_lagrange_dev", "zk_coeff_to_extended_dev", "zk_extended_to_coeff_dev",
    "zk_divide_by_vanishing_poly_dev",
    "zk_pk_upload", "zk_pk_destroy", "zk_pk_domain", "zk_pk_program_stats", "zk_evalh_generate_source", "zk_pk_evaluator_kind", "zk_pk_evaluator_log",
    "zk_evaluate_h", "zk_evaluate_h_dev",
    "zk_keygen", "zk_pkey_destroy", "zk_pkey_vk_commitments", "zk_pkey_pk", "zk_pkey_aux_ctx", "zk_pkey_set_transcript", "zk_proof_randoms", "zk_proof_max_bytes",
    "zk_create_proof", "zk_verify_proof", "zk_g2_decode", "zk_g2_encode", "zk_poseidon_permute", "zk_poseidon_hash",
    "zk_group_unique_id", "zk_group_create", "zk_group_destroy", "zk_group_rank", "zk_group_world", "zk_group_p2p_enabled", "zk_pkey_set_group", "zk_msm_sharded_dev",
    "zk_group_shard_plan", "zk_group_balance_plan",
    "zk_field_op", "zk_bench_modmul", "zk_bench_imad",
]

_lib = None


def load():
    """Load the CUDA library.  There is no fallback: a missing library is an error."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ZkError("%s is missing: run `python zk-dcap-verifier_b200/build.py` (or __graft_entry__.build()); "
                      "this backend has no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, sz, u32, i32 = C.c_void_p, C.c_size_t, C.c_uint32, C.c_int
    protos = {
        "zk_ctx_create": (i32, [i32, vp, C.POINTER(vp)]),
        "zk_ctx_destroy": (None, [vp]),
        "zk_last_error": (C.c_char_p, [vp]),
        "zk_ctx_sync": (i32, [vp]),
        "zk_version": (C.c_char_p, []),
        "zk_ctx_timing": (i32, [vp, i32]),
        "zk_ctx_timing_reset": (i32, [vp]),
        "zk_ctx_timing_read": (i32, [vp, C.POINTER(zk_kernel_stat), i32]),
        "zk_ctx_launch_count": (C.c_uint64, [vp]),
        "zk_dev_alloc": (i32, [vp, sz, C.POINTER(vp)]),
        "zk_dev_free": (i32, [vp, vp]),
        "zk_h2d": (i32, [vp, vp, vp, sz]),
        "zk_d2h": (i32, [vp, vp, vp, sz]),
        "zk_srs_upload": (i32, [vp, u32, vp, vp, i32]),
        "zk_srs_upload_dev": (i32, [vp, u32, vp, vp, i32]),
        "zk_g1_fixed_base_mul_dev": (i32, [vp, vp, vp, sz, vp]),
        "zk_ctx_share_srs": (i32, [vp, vp]),
        "zk_srs_window": (i32, [vp, i32, C.POINTER(u32), C.POINTER(u32)]),
        "zk_ctx_counters": (i32, [vp, C.POINTER(C.c_uint64), i32]),
        "zk_params_read": (i32, [vp, vp, sz, i32, C.POINTER(u32), vp, vp]),
        "zk_params_write": (i32, [vp, vp, vp, vp, sz, C.POINTER(sz)]),
        "zk_g1_compress": (i32, [vp, vp, sz, vp]),
        "zk_g1_decompress": (i32, [vp, vp, sz, vp]),
        "zk_g1_encode": (i32, [vp, i32, vp, sz, vp]),
        "zk_g1_decode": (i32, [vp, i32, vp, sz, vp]),
        "zk_params_read_format": (i32, [vp, i32, vp, sz, i32, C.POINTER(u32), vp, vp]),
        "zk_params_write_format": (i32, [vp, i32, vp, vp, vp, sz, C.POINTER(sz)]),
        "zk_msm": (i32, [vp, vp, vp, i32, sz, vp]),
        "zk_msm_dev": (i32, [vp, vp, vp, i32, sz, vp]),
        "zk_msm_batch_dev": (i32, [vp, vp, i32, sz, sz, vp]),
        "zk_g1_sum": (i32, [vp, vp, sz, vp]),
        "zk_g1_sum_dev": (i32, [vp, vp, sz, vp]),
        "zk_ntt": (i32, [vp, vp, vp, u32]),
        "zk_ntt_dev": (i32, [vp, vp, vp, u32]),
        "zk_domain_create": (i32, [vp, u32, u32, C.POINTER(vp)]),
        "zk_domain_destroy": (None, [vp]),
        "zk_domain_k": (u32, [vp]),
        "zk_domain_extended_k": (u32, [vp]),
        "zk_domain_constants": (i32, [vp, vp]),
        "zk_lagrange_to_coeff": (i32, [vp, vp]),
        "zk_coeff_to_lagrange": (i32, [vp, vp]),
        "zk_coeff_to_extended": (i32, [vp, vp, vp]),
        "zk_extended_to_coeff": (i32, [vp, vp, vp]),
        "zk_divide_by_vanishing_poly": (i32, [vp, vp]),
        "zk_lagrange_to_coeff_dev": (i32, [vp, vp, sz]),
        "zk_coeff_to_lagrange_dev": (i32, [vp, vp, sz]),
        "zk_coeff_to_extended_dev": (i32, [vp, vp, vp, sz]),
        "zk_extended_to_coeff_dev": (i32, [vp, vp, vp]),
        "zk_divide_by_vanishing_poly_dev": (i32, [vp, vp]),
        "zk_pk_upload": (i32, [vp, C.POINTER(zk_evaluator_desc), vp, vp, vp, vp, vp, i32, C.POINTER(vp)]),
        "zk_pk_destroy": (None, [vp]),
        "zk_pk_domain": (vp, [vp]),
        "zk_pk_program_stats": (i32, [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
        "zk_evalh_generate_source": (i32, [C.POINTER(zk_evaluator_desc), i32, vp, sz, C.POINTER(sz)]),
        "zk_pk_evaluator_kind": (i32, [vp]),
        "zk_pk_evaluator_log": (C.c_char_p, [vp, C.POINTER(C.c_double)]),
        "zk_evaluate_h": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "zk_evaluate_h_dev": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "zk_keygen": (i32, [vp, C.POINTER(zk_circuit_desc), vp, vp, C.POINTER(vp)]),
        "zk_pkey_destroy": (None, [vp]),
        "zk_pkey_vk_commitments": (i32, [vp, vp, vp]),
        "zk_pkey_pk": (vp, [vp]),
        "zk_pkey_aux_ctx": (vp, [vp]),
        "zk_pkey_set_transcript": (i32, [vp, i32]),
        "zk_proof_randoms": (sz, [vp]),
        "zk_proof_max_bytes": (sz, [vp]),
        "zk_create_proof": (i32, [vp, vp, vp, vp, vp, sz, vp, sz, C.POINTER(sz)]),
        "zk_verify_proof": (i32, [C.POINTER(zk_circuit_desc), C.POINTER(zk_verifier_params), vp, vp, vp, sz, C.POINTER(i32)]),
        "zk_poseidon_permute": (i32, [vp, u32, u32, u32]),
        "zk_poseidon_hash": (i32, [vp, sz, vp, u32]),
        "zk_g2_decode": (i32, [i32, vp, vp]),
        "zk_g2_encode": (i32, [i32, vp, vp]),
        "zk_group_unique_id": (i32, [vp]),
        "zk_group_create": (i32, [vp, u32, u32, vp, C.POINTER(vp)]),
        "zk_group_destroy": (None, [vp]),
        "zk_group_rank": (u32, [vp]),
        "zk_group_world": (u32, [vp]),
        "zk_group_p2p_enabled": (i32, [vp]),
        "zk_pkey_set_group": (i32, [vp, vp]),
        "zk_msm_sharded_dev": (i32, [vp, vp, i32, sz, vp]),
        "zk_group_balance_plan": (None, [C.POINTER(u32), sz, u32, C.POINTER(u32)]),
        "zk_group_shard_plan": (None, [u32, u32, sz, sz, C.POINTER(C.c_int), C.POINTER(sz), C.POINTER(sz), C.POINTER(sz)]),
        "zk_field_op": (i32, [vp, i32, i32, vp, vp, vp, sz]),
        "zk_bench_modmul": (i32, [vp, i32, u32, C.POINTER(C.c_double)]),
        "zk_bench_imad": (i32, [vp, i32, u32, C.POINTER(C.c_double)]),
    }
    for name in SYMBOLS:
        fn = getattr(lib, name)          # raises AttributeError if the export is missing
        fn.restype, fn.argtypes = protos[name]
    _lib = lib
    return lib
