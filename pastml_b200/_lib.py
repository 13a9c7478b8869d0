"""ctypes binding of ``libpastml_b200.so`` (the C ABI of ``include/pastml_b200.h``).

The product path has no CPU fallback: importing this module never builds or computes anything, but
:func:`load` raises if the CUDA engine has not been built, and the engine itself refuses to create a problem
without a CUDA device.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('PASTML_B200_LIB') or os.path.join(HERE, 'libpastml_b200.so')

EXPORTS = ['pml_last_error', 'pml_version', 'pml_device_count', 'pml_param_stride', 'pml_problem_create',
           'pml_problem_destroy', 'pml_loglik', 'pml_joint', 'pml_marginal', 'pml_mppa_finalize', 'pml_parsimony',
           'pml_set_custom_masks', 'pml_get_masks', 'pml_get_initial', 'pml_zero_likelihood_info',
           'pml_zero_likelihood_all', 'pml_trim_memory', 'pml_get_root_partials', 'pml_set_tip_partials',
           'pml_get_tip_td', 'pml_set_root_td',
           'pml_last_call_stats', 'pml_last_kernels', 'pml_marginal_counts', 'pml_bench_pass', 'pml_supports_lane_lists', 'pml_loglik_lanes', 'pml_get_bottom_up',
           'pml_shard_create', 'pml_shard_destroy', 'pml_shard_unique_id', 'pml_shard_connect', 'pml_shard_bottom_up',
           'pml_shard_exchange_local', 'pml_shard_finish', 'pml_shard_top_down', 'pml_shard_results',
           'pml_shard_joint_bottom_up', 'pml_shard_joint_finish', 'pml_shard_joint_states',
           'pml_shard_bench_step',
           'pml_newick_last_error', 'pml_newick_parse', 'pml_newick_num_nodes', 'pml_newick_num_trees',
           'pml_newick_names_bytes', 'pml_newick_name_nodes', 'pml_newick_fill', 'pml_newick_free',
           'pml_write_float_table']

PML_OK, PML_ERR_INVALID, PML_ERR_CUDA, PML_ERR_ZERO_LIKELIHOOD, PML_ERR_NO_DEVICE = 0, 1, 2, 3, 4
MODEL_F81, MODEL_HKY, MODEL_EIG = 0, 1, 2
MASK_ANNOTATION, MASK_JOINT, MASK_MAP, MASK_MPPA, MASK_CUSTOM = 0, 1, 2, 3, 4
PARS_ACCTRAN, PARS_DOWNPASS, PARS_DELTRAN = 1, 2, 4

_lib = None


class EngineNotBuilt(RuntimeError):
    pass


def load():
    """Loads the engine library (built by ``python -m pastml_b200.build`` / ``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineNotBuilt('{} is missing: build the CUDA engine first (python pastml_b200/build.py). '
                             'pastml_b200 has no CPU fallback.'.format(LIB_PATH))
    lib = C.CDLL(LIB_PATH)
    vp, i32, dbl = C.c_void_p, C.c_int32, C.c_double
    lib.pml_last_error.restype = C.c_char_p
    lib.pml_last_error.argtypes = []
    lib.pml_version.restype = C.c_int
    lib.pml_device_count.restype = C.c_int
    lib.pml_param_stride.restype = C.c_int
    lib.pml_param_stride.argtypes = [i32, i32]
    lib.pml_problem_create.restype = C.c_int
    lib.pml_problem_create.argtypes = [i32, vp, vp, i32, i32, i32, vp, vp, i32, i32, C.POINTER(vp)]
    lib.pml_problem_destroy.restype = None
    lib.pml_problem_destroy.argtypes = [vp]
    lib.pml_shard_create.restype = C.c_int
    lib.pml_shard_create.argtypes = [vp, vp, i32, i32, i32, vp, vp, C.POINTER(vp)]
    lib.pml_shard_destroy.restype = None
    lib.pml_shard_destroy.argtypes = [vp]
    lib.pml_shard_unique_id.restype = C.c_int
    lib.pml_shard_unique_id.argtypes = [vp]
    lib.pml_shard_connect.restype = C.c_int
    lib.pml_shard_connect.argtypes = [vp, vp]
    lib.pml_shard_bottom_up.restype = C.c_int
    lib.pml_shard_bottom_up.argtypes = [vp, vp, i32, i32]
    lib.pml_shard_exchange_local.restype = C.c_int
    lib.pml_shard_exchange_local.argtypes = [vp, i32]
    lib.pml_shard_finish.restype = C.c_int
    lib.pml_shard_finish.argtypes = [vp, vp, i32, vp]
    lib.pml_shard_top_down.restype = C.c_int
    lib.pml_shard_top_down.argtypes = [vp, i32, i32, i32]
    lib.pml_shard_joint_bottom_up.restype = C.c_int
    lib.pml_shard_joint_bottom_up.argtypes = [vp, vp]
    lib.pml_shard_joint_finish.restype = C.c_int
    lib.pml_shard_joint_finish.argtypes = [vp, vp, vp]
    lib.pml_shard_joint_states.restype = C.c_int
    lib.pml_shard_joint_states.argtypes = [vp, i32, vp]
    lib.pml_shard_results.restype = C.c_int
    lib.pml_shard_results.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    lib.pml_shard_bench_step.restype = C.c_int
    lib.pml_shard_bench_step.argtypes = [vp, vp, i32, i32, vp, vp]
    lib.pml_loglik.restype = C.c_int
    lib.pml_loglik.argtypes = [vp, vp, i32, i32, i32, i32, vp]
    lib.pml_supports_lane_lists.restype = C.c_int
    lib.pml_supports_lane_lists.argtypes = [vp]
    lib.pml_loglik_lanes.restype = C.c_int
    lib.pml_loglik_lanes.argtypes = [vp, vp, vp, i32, vp]
    lib.pml_get_bottom_up.restype = C.c_int
    lib.pml_get_bottom_up.argtypes = [vp, vp, vp, vp, vp]
    lib.pml_joint.restype = C.c_int
    lib.pml_joint.argtypes = [vp, vp, vp, vp]
    lib.pml_marginal.restype = C.c_int
    lib.pml_marginal.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp, vp]
    lib.pml_mppa_finalize.restype = C.c_int
    lib.pml_mppa_finalize.argtypes = [vp, i32, vp, vp]
    lib.pml_parsimony.restype = C.c_int
    lib.pml_parsimony.argtypes = [vp, i32, vp, vp]
    lib.pml_set_custom_masks.restype = C.c_int
    lib.pml_set_custom_masks.argtypes = [vp, vp]
    lib.pml_get_masks.restype = C.c_int
    lib.pml_get_masks.argtypes = [vp, i32, vp]
    lib.pml_get_initial.restype = C.c_int
    lib.pml_get_initial.argtypes = [vp, vp, vp]
    lib.pml_zero_likelihood_info.restype = C.c_int
    lib.pml_zero_likelihood_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    lib.pml_zero_likelihood_all.restype = C.c_int
    lib.pml_zero_likelihood_all.argtypes = [vp, vp, i32]
    lib.pml_get_root_partials.restype = C.c_int
    lib.pml_get_root_partials.argtypes = [vp, i32, vp, vp]
    lib.pml_set_tip_partials.restype = C.c_int
    lib.pml_set_tip_partials.argtypes = [vp, i32, vp, vp, vp, i32]
    i64 = C.c_int64
    lib.pml_newick_last_error.restype = C.c_char_p
    lib.pml_newick_last_error.argtypes = []
    lib.pml_newick_parse.restype = C.c_int
    lib.pml_newick_parse.argtypes = [C.c_char_p, i64, C.POINTER(vp)]
    lib.pml_newick_num_nodes.restype = i64
    lib.pml_newick_num_nodes.argtypes = [vp]
    lib.pml_newick_num_trees.restype = i64
    lib.pml_newick_num_trees.argtypes = [vp]
    lib.pml_newick_names_bytes.restype = i64
    lib.pml_newick_names_bytes.argtypes = [vp]
    lib.pml_write_float_table.restype = C.c_int
    lib.pml_write_float_table.argtypes = [C.c_char_p, C.c_char_p, vp, vp, vp, C.c_int64, C.c_int32]
    lib.pml_newick_name_nodes.restype = C.c_int
    lib.pml_newick_name_nodes.argtypes = [vp, C.c_int]
    lib.pml_newick_fill.restype = C.c_int
    lib.pml_newick_fill.argtypes = [vp, vp, vp, vp, vp]
    lib.pml_newick_free.restype = None
    lib.pml_newick_free.argtypes = [vp]
    lib.pml_get_tip_td.restype = C.c_int
    lib.pml_get_tip_td.argtypes = [vp, vp, vp]
    lib.pml_set_root_td.restype = C.c_int
    lib.pml_set_root_td.argtypes = [vp, vp, vp]
    lib.pml_trim_memory.restype = C.c_int
    lib.pml_trim_memory.argtypes = []
    lib.pml_marginal_counts.restype = C.c_int
    lib.pml_marginal_counts.argtypes = [vp, vp, C.c_int32, C.c_uint64, vp]
    lib.pml_last_kernels.restype = C.c_int
    lib.pml_last_kernels.argtypes = [vp, C.c_char_p, C.c_int32]
    lib.pml_last_call_stats.restype = C.c_int
    lib.pml_last_call_stats.argtypes = [vp, C.POINTER(dbl), C.POINTER(C.c_int64)]
    lib.pml_bench_pass.restype = C.c_int
    lib.pml_bench_pass.argtypes = [vp, vp, i32, i32, i32, i32, C.POINTER(dbl), C.POINTER(dbl), C.POINTER(C.c_int64)]
    _lib = lib
    return lib


def last_error():
    return load().pml_last_error().decode('utf-8', 'replace')
