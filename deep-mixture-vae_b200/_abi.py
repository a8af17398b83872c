"""ctypes binding of libdmvae_b200.so (include/dmvae_b200.h).

There is no CPU fallback: if the shared library is missing (and cannot be built because nvcc is absent) the
import fails loudly, and every compute entry point needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _build

_lib = None

# enums (include/dmvae_b200.h)
F32, BF16, U8 = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_SIGMOID = 0, 1, 2
INPUT_BINARY, INPUT_REAL = 0, 1
MODE_DMVAE, MODE_DMVAE_SAMPLED, MODE_VADE = 0, 1, 2

c_void_p, c_int, c_int32, c_int64, c_float, c_uint64 = C.c_void_p, C.c_int, C.c_int32, C.c_int64, C.c_float, C.c_uint64


class DmvaeError(RuntimeError):
    pass


class ReconFuse(C.Structure):
    _fields_ = [("X", c_void_p), ("x_dtype", c_int32), ("ldx", c_int64), ("x_scale", c_float), ("input_type", c_int32),
                ("scale", c_float), ("D", c_int32), ("r_part", c_void_p), ("r_parts", c_int32)]


class GemmEpilogue(C.Structure):
    _fields_ = [("out_dtype", c_int32), ("act", c_int32), ("n_valid", c_int32), ("n_block", c_int32),
                ("pad_one", c_float), ("relu_mask", c_void_p), ("ld_mask", c_int64), ("bias", c_void_p),
                ("accumulate", c_int32), ("split_k", c_int32), ("recon", C.POINTER(ReconFuse))]


class ChainGemm(C.Structure):
    _fields_ = [("trans_a", c_int32), ("trans_b", c_int32), ("A", c_void_p), ("lda", c_int64), ("B", c_void_p),
                ("ldb", c_int64), ("C", c_void_p), ("ldc", c_int64), ("M", c_int32), ("N", c_int32), ("K", c_int32),
                ("epi", GemmEpilogue), ("dep", c_int32 * 2), ("dep_all", c_int32 * 2), ("fuse", c_int32)]


class ReparamArgs(C.Structure):
    _fields_ = [("rows", c_int32), ("L", c_int32), ("K", c_int32),
                ("mean", c_void_p), ("log_var", c_void_p), ("ld_zh", c_int64),
                ("logits", c_void_p), ("ld_logits", c_int64),
                ("eps_in", c_void_p), ("gumbel_in", c_void_p),
                ("seed", c_uint64), ("step", c_uint64), ("row_offset", c_uint64), ("step_dev", c_void_p),
                ("tau", c_float),
                ("Z_out", c_void_p), ("z_dtype", c_int32), ("ld_z", c_int64), ("z_cols", c_int32),
                ("eps_out", c_void_p), ("zeta_out", c_void_p),
                ("fold", c_void_p), ("ld_fold", c_int64), ("fold_K", c_int32), ("fold_stride", c_int32)]


class ElboArgs(C.Structure):
    _fields_ = [("mode", c_int32), ("input_type", c_int32),
                ("rows", c_int32), ("D", c_int32), ("L", c_int32), ("K", c_int32),
                ("X", c_void_p), ("x_dtype", c_int32), ("ldx", c_int64),
                ("decoded", c_void_p), ("dec_dtype", c_int32), ("ld_dec", c_int64),
                ("mean", c_void_p), ("log_var", c_void_p), ("ld_zh", c_int64),
                ("logits", c_void_p), ("ld_logits", c_int64),
                ("eps", c_void_p), ("ld_eps", c_int64),
                ("zeta", c_void_p), ("ld_zeta", c_int64),
                ("tau", c_float),
                ("prior_means", c_void_p), ("prior_log_vars", c_void_p),
                ("kl_ratio", c_float), ("inv_global_batch", c_float), ("kl_ratio_dev", c_void_p), ("recon_scale", c_float),
                ("per_sample", c_void_p), ("qc", c_void_p), ("argmax", c_void_p),
                ("d_decoded", c_void_p), ("ld_ddec", c_int64), ("ddec_cols", c_int32),
                ("d_mean_kl", c_void_p), ("d_log_var_kl", c_void_p), ("ld_dkl", c_int64),
                ("d_logits", c_void_p), ("dlogits_dtype", c_int32), ("ld_dlogits", c_int64), ("dlogits_cols", c_int32),
                ("d_Z_gamma", c_void_p), ("ld_dzg", c_int64),
                ("w_scratch", c_void_p), ("f_scratch", c_void_p),
                ("d_gate_extra", c_void_p), ("ld_dge", c_int64), ("x_scale", c_float),
                ("r_part", c_void_p), ("r_parts", c_int32)]


class MoeArgs(C.Structure):
    _fields_ = [("classification", c_int32), ("rows", c_int32), ("E", c_int32), ("O", c_int32),
                ("pred", c_void_p), ("ld_pred", c_int64),
                ("gate", c_void_p), ("ld_gate", c_int64),
                ("Y", c_void_p), ("ldy", c_int64),
                ("inv_global_batch", c_float),
                ("per_sample", c_void_p), ("y_soft", c_void_p), ("pred_class", c_void_p),
                ("d_pred", c_void_p), ("dpred_dtype", c_int32), ("ld_dpred", c_int64), ("dpred_cols", c_int32),
                ("d_gate", c_void_p), ("ld_dgate", c_int64)]


# name -> (restype, argtypes); every symbol include/dmvae_b200.h declares
SIGNATURES = {
    "dmvae_abi_version": (c_int, []),
    "dmvae_last_error": (C.c_char_p, []),
    "dmvae_ctx_create": (c_int, [c_int, C.POINTER(c_void_p)]),
    "dmvae_ctx_destroy": (c_int, [c_void_p]),
    "dmvae_ctx_launch_count": (c_int64, [c_void_p]),
    "dmvae_ctx_has_tcgen05": (c_int, [c_void_p]),
    "dmvae_gemm": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64,
                           c_int, c_int, c_int, C.POINTER(GemmEpilogue), c_void_p]),
    "dmvae_gemm_chain_counters": (c_int64, [c_int, c_int]),
    "dmvae_gemm_chain": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "dmvae_linear_fwd": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int,
                                 c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dmvae_linear_dgrad": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64,
                                   c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dmvae_linear_wgrad": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64,
                                   c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "dmvae_stage_input": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_void_p, c_int, c_int64, c_int, c_int, c_float, c_void_p]),
    "dmvae_gather_rows": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "dmvae_gather_rows_bits": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "dmvae_reparam_fwd": (c_int, [c_void_p, C.POINTER(ReparamArgs), c_void_p]),
    "dmvae_reparam_bwd": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p,
                                  c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int, c_int64, c_int, c_void_p]),
    "dmvae_elbo_fwd_bwd": (c_int, [c_void_p, C.POINTER(ElboArgs), c_void_p]),
    "dmvae_elbo_reduce_workspace": (c_int64, [c_int, c_int, c_int]),
    "dmvae_elbo_reduce": (c_int, [c_void_p, C.POINTER(ElboArgs), c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "dmvae_elbo_reduce_stage": (c_int, [c_void_p, C.POINTER(ElboArgs), c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int,
                                        c_void_p]),
    "dmvae_moe_fwd_bwd": (c_int, [c_void_p, C.POINTER(MoeArgs), c_void_p]),
    "dmvae_softmax_bwd_add": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_int, c_int64,
                                      c_int, c_int, c_void_p]),
    "dmvae_softmax_rows": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p]),
    "dmvae_reduce_columns": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_float, c_void_p, c_void_p]),
    "dmvae_stage_features": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_int, c_int64, c_int, c_void_p]),
    "dmvae_adam": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_void_p, c_float,
                           c_float, c_float, c_float, c_int, c_void_p]),
    "dmvae_step_tick": (c_int, [c_void_p, c_void_p, c_float, c_float, c_float, c_void_p]),
    "dmvae_log_append": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, C.c_uint64, c_void_p]),
    "dmvae_argmax_contingency": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_int, c_void_p,
                                         c_void_p, c_void_p]),
    "dmvae_dp_reduce_adam": (c_int, [c_void_p, c_int, c_int, C.POINTER(c_void_p), C.POINTER(c_void_p),
                                     C.POINTER(c_void_p), C.POINTER(c_void_p), C.POINTER(c_int64), c_int, c_void_p, c_void_p,
                                     c_int64, c_int64, c_int64, c_float, c_void_p, c_float, c_float, c_float, c_int, c_void_p]),
    "dmvae_dp_reduce_adam_mc": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                                        c_int64, c_float, c_void_p, c_float, c_float, c_float, c_void_p]),
    "dmvae_dp_alloc": (c_int, [c_void_p, c_int64, C.POINTER(c_void_p), c_void_p]),
    "dmvae_dp_open": (c_int, [c_void_p, c_void_p, C.POINTER(c_void_p)]),
    "dmvae_dp_close": (c_int, [c_void_p, c_void_p]),
    "dmvae_dp_free": (c_int, [c_void_p, c_void_p]),
    "dmvae_dp_barrier": (c_int, [c_void_p, c_int, c_int, C.POINTER(c_void_p), c_void_p, c_int, c_void_p]),
    "dmvae_zero_f32": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "dmvae_cast_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "dmvae_split3_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_void_p]),
    "dmvae_fold3": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p]),
}


def library_path() -> str:
    return _build.LIB


def load():
    """Load (building first if the sources are newer and nvcc is available) and type every entry point."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if not os.path.exists(path) or (os.path.exists(_build.NVCC) and os.environ.get("DMVAE_B200_NO_REBUILD") != "1"
                                    and _build._stale()):
        if not os.path.exists(_build.NVCC):
            raise ImportError("libdmvae_b200.so is missing at %s and nvcc is not available to build it; "
                              "there is no CPU fallback" % path)
        _build.build()
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.dmvae_abi_version() != 2:
        raise ImportError("libdmvae_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().dmvae_last_error()
        raise DmvaeError("dmvae_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))
