"""Eager entry points of the prior maths on CUDA tensors, each one a call of the fused kernels through the
C ABI (used by priors.py so that ``kl_from_prior`` / ``inverse_reparametrize`` / ``get_cluster_probs`` work
outside a model, and by the tests).  No torch arithmetic: tensors only carry memory."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _abi

_ctx_cache = {}


def _ctx(device: torch.device):
    lib = _abi.load()
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _ctx_cache:
        c = C.c_void_p()
        _abi.check(lib.dmvae_ctx_create(idx, C.byref(c)))
        _ctx_cache[idx] = c
    return lib, _ctx_cache[idx]


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _f32(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError("dmvae_b200.functional needs CUDA tensors: there is no CPU fallback")
    return t.contiguous().to(torch.float32) if (t.dtype != torch.float32 or not t.is_contiguous()) else t


def elbo_terms(mode: int, mean, log_var, prior_means, prior_log_vars, logits=None, eps=None, zeta=None, tau=1.0,
               X=None, decoded=None, input_type="real", kl_ratio=1.0, inv_global_batch=None):
    """Run the fused ELBO kernel; returns dict(per_sample [B,4], qc [B,K], argmax [B], d_mean, d_log_var, d_logits,
    d_Z_gamma, d_decoded).  With X/decoded omitted the reconstruction term is an 8-column zero dummy."""
    mean, log_var = _f32(mean), _f32(log_var)
    dev = mean.device
    lib, ctx = _ctx(dev)
    B, L = mean.shape
    pm, pl = _f32(prior_means), _f32(prior_log_vars)
    K = pm.shape[0]
    if X is None:
        X = torch.zeros(B, 8, device=dev)
        decoded = torch.zeros(B, 8, device=dev)
    X, decoded = _f32(X), _f32(decoded)
    D = X.shape[1]
    zh = torch.empty(B, 2 * L, device=dev)
    zh[:, :L].copy_(mean)
    zh[:, L:].copy_(log_var)
    ea = _abi.ElboArgs()
    ea.mode, ea.input_type = mode, (_abi.INPUT_BINARY if input_type == "binary" else _abi.INPUT_REAL)
    ea.rows, ea.D, ea.L, ea.K = B, D, L, K
    ea.X, ea.x_dtype, ea.ldx = X.data_ptr(), _abi.F32, X.stride(0)
    ea.decoded, ea.dec_dtype, ea.ld_dec = decoded.data_ptr(), _abi.F32, decoded.stride(0)
    ea.mean, ea.log_var, ea.ld_zh = zh.data_ptr(), zh.data_ptr() + 4 * L, 2 * L
    out = dict(per_sample=torch.empty(B, 4, device=dev), qc=torch.empty(B, K, device=dev),
               argmax=torch.empty(B, dtype=torch.int32, device=dev), d_decoded=torch.empty(B, D, device=dev),
               d_mean=torch.empty(B, L, device=dev), d_log_var=torch.empty(B, L, device=dev),
               d_logits=torch.zeros(B, K, device=dev), d_Z_gamma=torch.zeros(B, L, device=dev))
    keep = [zh, pm, pl, X, decoded]
    if logits is not None:
        lg = _f32(logits)
        keep.append(lg)
        ea.logits, ea.ld_logits = lg.data_ptr(), lg.stride(0)
    ea.d_logits, ea.dlogits_dtype, ea.ld_dlogits, ea.dlogits_cols = out["d_logits"].data_ptr(), _abi.F32, K, K
    if eps is not None:
        e = _f32(eps)
        keep.append(e)
        ea.eps, ea.ld_eps = e.data_ptr(), e.stride(0)
    if zeta is not None:
        z = _f32(zeta)
        keep.append(z)
        ea.zeta, ea.ld_zeta = z.data_ptr(), z.stride(0)
    ea.tau = tau
    ea.prior_means, ea.prior_log_vars = pm.data_ptr(), pl.data_ptr()
    ea.kl_ratio, ea.inv_global_batch, ea.recon_scale = kl_ratio, (1.0 / B if inv_global_batch is None else inv_global_batch), 1.0
    ea.per_sample, ea.qc, ea.argmax = out["per_sample"].data_ptr(), out["qc"].data_ptr(), out["argmax"].data_ptr()
    ea.d_decoded, ea.ld_ddec, ea.ddec_cols = out["d_decoded"].data_ptr(), D, D
    ea.d_mean_kl, ea.d_log_var_kl, ea.ld_dkl = out["d_mean"].data_ptr(), out["d_log_var"].data_ptr(), L
    ea.d_Z_gamma, ea.ld_dzg = out["d_Z_gamma"].data_ptr(), L
    ws, fs = torch.empty(B, K, device=dev), torch.empty(B, 2 * L, device=dev)
    ea.w_scratch, ea.f_scratch = ws.data_ptr(), fs.data_ptr()
    _abi.check(lib.dmvae_elbo_fwd_bwd(ctx, C.byref(ea), _stream(dev)))
    return out


def reparametrize(mean, log_var, epsilon=None, logits=None, gumbel=None, tau=1.0, seed=0, step=0):
    """Z = mean + exp(log_var/2) * eps (priors.py:86-89); with logits also zeta = softmax((logits+g)/tau)
    (priors.py:170-181).  eps / gumbel None -> Philox.  Returns (Z [B,L], eps_used, zeta or None)."""
    mean, log_var = _f32(mean), _f32(log_var)
    dev = mean.device
    lib, ctx = _ctx(dev)
    B, L = mean.shape
    zh = torch.empty(B, 2 * L, device=dev)
    zh[:, :L].copy_(mean)
    zh[:, L:].copy_(log_var)
    ra = _abi.ReparamArgs()
    K = logits.shape[-1] if logits is not None else 0
    ra.rows, ra.L, ra.K = B, L, K
    ra.mean, ra.log_var, ra.ld_zh = zh.data_ptr(), zh.data_ptr() + 4 * L, 2 * L
    keep = [zh]
    if logits is not None:
        lg = _f32(logits.reshape(B, K))
        keep.append(lg)
        ra.logits, ra.ld_logits = lg.data_ptr(), K
    if epsilon is not None:
        e = _f32(epsilon)
        keep.append(e)
        ra.eps_in = e.data_ptr()
    if gumbel is not None:
        g = _f32(gumbel.reshape(B, K))
        keep.append(g)
        ra.gumbel_in = g.data_ptr()
    ra.seed, ra.step, ra.row_offset, ra.tau = seed, step, 0, tau
    Z = torch.empty(B, L, device=dev)
    eo = torch.empty(B, L, device=dev)
    zeta = torch.empty(B, K, device=dev) if logits is not None else None
    ra.Z_out, ra.z_dtype, ra.ld_z, ra.z_cols = Z.data_ptr(), _abi.F32, L, L
    ra.eps_out = eo.data_ptr()
    ra.zeta_out = zeta.data_ptr() if zeta is not None else None
    _abi.check(lib.dmvae_reparam_fwd(ctx, C.byref(ra), _stream(dev)))
    return Z, eo, zeta
