"""ctypes binding of libqsb200.so -- exactly the entry points declared in include/qsb200.h.

There is no CPU fallback: if the CUDA library is missing the import of any compute entry
point raises.  (`python -m quicksand_b200.build` compiles it in-tree with nvcc.)
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QSB200_LIB") or os.path.join(HERE, "libqsb200.so")   # QSB200_LIB: another build of the same ABI (A/B measurements)

FAM_INDEP, FAM_GAUSS_PREC, FAM_LOGISTIC, FAM_EIGHT_SCHOOLS, FAM_FUNNEL = 1, 2, 3, 4, 5
ERP_GAUSSIAN, ERP_GAMMA, ERP_BETA, ERP_UNIFORM, ERP_DIRICHLET, ERP_NEGGAMMA = 0, 1, 2, 3, 4, 5
KERNEL_HMC, KERNEL_DRIFT, KERNEL_HARM, KERNEL_TRACEMH = 1, 2, 3, 4
FLAG_MOMENTS, FLAG_RECORD, FLAG_RECORD_LEAP, FLAG_TIME_KERNEL, FLAG_BALANCED, FLAG_INVARIANT = 1, 2, 4, 8, 16, 32

COMM_ID_BYTES = 128
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int32)


class ModelDesc(C.Structure):
    _fields_ = [("family", C.c_int32), ("dim", C.c_int32), ("nobs", C.c_int32), ("ret_mode", C.c_int32),
                ("ret_div", C.c_double), ("prior_sigma", C.c_double),
                ("erp_kind", ip), ("erp_p0", dp), ("erp_p1", dp), ("A", dp), ("X", dp),
                ("ybin", C.POINTER(C.c_uint8)), ("y", dp), ("sigma", dp), ("erp_lexid", ip), ("erp_fmu", dp), ("erp_fsigma", dp),
                ("erp_pref", ip), ("erp_pcoef", dp), ("erp_ptf", ip), ("erp_cond_lo", dp), ("erp_cond_hi", dp)]


class KernelSpec(C.Structure):
    _fields_ = [("kind", C.c_int32), ("doAdapt", C.c_int32), ("stepSize", C.c_double),
                ("numSteps", C.c_uint64), ("scale", C.c_double), ("weight", C.c_double),
                ("lexicalScaleSharing", C.c_int32), ("reserved", C.c_int32)]


class RunConfig(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("chain_begin", C.c_uint32), ("numchains", C.c_uint32),
                ("sample_chains", C.c_uint32), ("flags", C.c_uint32), ("device", C.c_int32),
                ("reserved", C.c_int32), ("stream", C.c_void_p)]


class Stats(C.Structure):
    _fields_ = [("props_made", C.c_uint64), ("props_accepted", C.c_uint64),
                ("sub_made", C.c_uint64 * 8), ("sub_accepted", C.c_uint64 * 8),
                ("grad_evals", C.c_uint64), ("leapfrog_steps", C.c_uint64), ("iterations", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("seconds", C.c_double), ("nonfinite", C.c_uint64), ("divergent", C.c_uint64)]


# name -> (restype, argtypes); every symbol include/qsb200.h declares
PROTOTYPES = {
    "qsb_version": (C.c_int, []),
    "qsb_last_error": (C.c_char_p, []),
    "qsb_source_hash": (C.c_char_p, []),
    "qsb_model_create": (C.c_int, [C.POINTER(ModelDesc), C.c_int32, C.POINTER(C.c_void_p)]),
    "qsb_model_destroy": (C.c_int, [C.c_void_p]),
    "qsb_model_update_data": (C.c_int, [C.c_void_p, C.POINTER(ModelDesc), C.c_void_p]),
    "qsb_model_dim": (C.c_int, [C.c_void_p]),
    "qsb_model_retdim": (C.c_int, [C.c_void_p]),
    "qsb_sampler_create": (C.c_int, [C.c_void_p, C.POINTER(KernelSpec), C.c_int32, C.POINTER(RunConfig), C.POINTER(C.c_void_p)]),
    "qsb_sampler_destroy": (C.c_int, [C.c_void_p]),
    "qsb_sampler_init": (C.c_int, [C.c_void_p]),
    "qsb_sampler_set_state": (C.c_int, [C.c_void_p, dp]),
    "qsb_sampler_get_state": (C.c_int, [C.c_void_p, dp]),
    "qsb_sampler_run": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64]),
    "qsb_sampler_begin": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64]),
    "qsb_sampler_advance": (C.c_int, [C.c_void_p, C.c_uint64]),
    "qsb_sampler_num_samples": (C.c_uint64, [C.c_void_p]),
    "qsb_sampler_set_temperatures": (C.c_int, [C.c_void_p, dp, C.c_uint64]),
    "qsb_sampler_fetch_samples": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_void_p, C.c_size_t]),
    "qsb_sampler_fetch_samples_async": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_void_p, C.c_size_t]),
    "qsb_sampler_fetch_sync": (C.c_int, [C.c_void_p]),
    "qsb_sampler_fetch_values": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, dp]),
    "qsb_sampler_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "qsb_sampler_kernel_time": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "qsb_sampler_set_kernel_timing": (C.c_int, [C.c_void_p, C.c_int32]),
    "qsb_sampler_diagnostics": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]),
    "qsb_sampler_fetch_record": (C.c_int, [C.c_void_p, dp, ip, dp, dp, ip, dp]),
    "qsb_query_expectation": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, dp, dp]),
    "qsb_query_expectation_host": (C.c_int, [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint64, C.c_int32, dp, dp]),
    "qsb_query_map": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, dp, dp]),
    "qsb_query_map_host": (C.c_int, [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint64, C.c_int32, dp, dp]),
    "qsb_query_autocorrelation": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, dp, C.c_double, dp]),
    "qsb_query_autocorrelation_host": (C.c_int, [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint64, C.c_int32, dp, C.c_double, dp]),
    "qsb_group_create": (C.c_int, [C.POINTER(ModelDesc), C.POINTER(KernelSpec), C.c_int32, C.POINTER(RunConfig), ip, C.c_int32, C.POINTER(C.c_void_p)]),
    "qsb_group_destroy": (C.c_int, [C.c_void_p]),
    "qsb_group_size": (C.c_int32, [C.c_void_p]),
    "qsb_group_sampler": (C.c_void_p, [C.c_void_p, C.c_int32]),
    "qsb_group_init": (C.c_int, [C.c_void_p]),
    "qsb_group_run": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64]),
    "qsb_group_begin": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64]),
    "qsb_group_advance": (C.c_int, [C.c_void_p, C.c_uint64]),
    "qsb_group_sync": (C.c_int, [C.c_void_p]),
    "qsb_group_fetch_samples": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_void_p, C.c_size_t]),
    "qsb_group_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "qsb_group_diagnostics": (C.c_int, [C.c_void_p, C.c_int32, dp]),
    "qsb_comm_unique_id": (C.c_int, [C.POINTER(C.c_uint8)]),
    "qsb_comm_create": (C.c_int, [C.POINTER(C.c_uint8), C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "qsb_comm_destroy": (C.c_int, [C.c_void_p]),
    "qsb_comm_allreduce_sum": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "qsb_sampler_diagnostics_allreduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, dp]),
    "qsb_measure_fp64_peak": (C.c_int, [C.c_int32, C.c_double, C.c_int32, C.POINTER(C.c_double)]),
    "qsb_debug_leapfrog": (C.c_int, [C.c_void_p, C.c_int32, dp, dp, dp, dp, C.c_int32, dp, dp, dp, dp]),
    "qsb_debug_glm_layout": (C.c_int, [C.c_uint32, C.c_int32, C.c_int32, C.c_int32, ip, ip, ip, ip, ip, ip]),
    "qsb_logp_grad": (C.c_int, [C.c_void_p, dp, C.c_int32, dp, dp, dp]),
    "qsb_logprob": (C.c_int, [C.c_int32, dp, C.c_int32, dp]),
    "qsb_uniforms": (C.c_int, [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int32, dp]),
    "qsb_gaussians": (C.c_int, [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int32, C.c_double, C.c_double, dp]),
}

_LIB = None


class QsbError(RuntimeError):
    """A non-zero return of the C ABI.  The Terra shim maps this to the reference's print + abort
    (qs/lib/std.t:35-43); Python raises."""


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libqsb200.so is not built (run `python -m quicksand_b200.build`); "
                              "quicksand_b200 has no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            f = getattr(L, name)   # raises AttributeError if the library does not export it
            f.restype = res
            f.argtypes = args
        _LIB = L
    return _LIB


def check(rc):
    if rc != 0:
        raise QsbError("qsb200 error %d: %s" % (rc, lib().qsb_last_error().decode()))
