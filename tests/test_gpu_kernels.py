"""GPU parity tests of the individual C-ABI kernels against the oracle (run on the B200 box).

Tolerances: fp32 kernels 1e-4 relative (north_star), bf16 tensor-core GEMMs 2e-2; integer outputs
(argmax, Philox stream, contingency counts) bit-exact."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import closed_form as cf
from oracle import philox
from oracle import reference_graph as rg


@pytest.fixture(scope="module")
def lib():
    from dmvae_b200 import _abi
    return _abi.load()


@pytest.fixture(scope="module")
def ctx(lib):
    from dmvae_b200 import _abi
    c = C.c_void_p()
    _abi.check(lib.dmvae_ctx_create(0, C.byref(c)))
    yield c
    lib.dmvae_ctx_destroy(c)


def dev(a, dtype=torch.float32):
    return torch.tensor(np.asarray(a), dtype=dtype, device="cuda")


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def relerr(got, ref):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    return float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30))


# ---------------------------------------------------------------------------------------------
# GEMM
# ---------------------------------------------------------------------------------------------
def _gemm(lib, ctx, dtype, ta, tb, A, B, M, N, K, out_dtype=0, **epi):
    from dmvae_b200 import _abi
    e = _abi.GemmEpilogue()
    e.out_dtype, e.act, e.n_valid, e.n_block = out_dtype, epi.get("act", 0), epi.get("n_valid", N), epi.get("n_block", N)
    e.pad_one = epi.get("pad_one", 1.0)
    mask = epi.get("mask")
    e.relu_mask = mask.data_ptr() if mask is not None else None
    e.ld_mask = mask.stride(0) if mask is not None else 0
    bias = epi.get("bias")
    e.bias = bias.data_ptr() if bias is not None else None
    e.accumulate, e.split_k = epi.get("accumulate", 0), epi.get("split_k", 1)
    Cm = epi.get("C")
    if Cm is None:
        Cm = torch.zeros(M, N, dtype=torch.float32 if out_dtype == 0 else torch.bfloat16, device="cuda")
    _abi.check(lib.dmvae_gemm(ctx, dtype, ta, tb, A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), Cm.data_ptr(),
                              Cm.stride(0), M, N, K, C.byref(e), stream()))
    torch.cuda.synchronize()
    return Cm


@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(100, 64, 832), (256, 192, 100), (37, 50, 23)])
def test_gemm_f32_matches_fp64(lib, ctx, ta, tb, M, N, K):
    rs = np.random.RandomState(0)
    a = rs.randn(M, K).astype(np.float32)
    b = rs.randn(K, N).astype(np.float32)
    A = dev(a.T.copy() if ta else a)
    B = dev(b.T.copy() if tb else b)
    got = _gemm(lib, ctx, 0, ta, tb, A, B, M, N, K).cpu().numpy()
    ref = a.astype(np.float64) @ b.astype(np.float64)
    assert relerr(got, ref) < 1e-5


def test_gemm_f32_epilogue_and_splitk(lib, ctx):
    rs = np.random.RandomState(1)
    M, N, K = 130, 128, 256
    a, b = rs.randn(M, K).astype(np.float32), rs.randn(K, N).astype(np.float32)
    mask = rs.randn(M, N).astype(np.float32)
    bias = rs.randn(N).astype(np.float32)
    ref = a.astype(np.float64) @ b + bias
    ref = np.maximum(ref, 0) * (mask > 0)
    j = np.arange(N) % 64
    ref[:, j == 50] = 1.0
    ref[:, j > 50] = 0.0
    got = _gemm(lib, ctx, 0, 0, 0, dev(a), dev(b), M, N, K, act=1, n_valid=50, n_block=64, mask=dev(mask),
                bias=dev(bias)).cpu().numpy()
    assert relerr(got, ref) < 1e-5
    Cm = torch.ones(M, N, device="cuda")
    got = _gemm(lib, ctx, 0, 0, 0, dev(a), dev(b), M, N, K, accumulate=1, split_k=4, C=Cm).cpu().numpy()
    assert relerr(got, 1.0 + a.astype(np.float64) @ b) < 1e-5


def _bf(x):
    return torch.tensor(x).to(torch.bfloat16).float().numpy()


@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(256, 128, 128), (4096, 512, 832), (100, 64, 64), (512, 832, 4096), (300, 2112, 512)])
def test_gemm_bf16_tcgen05(lib, ctx, ta, tb, M, N, K):
    rs = np.random.RandomState(2)
    a = _bf(rs.randn(M, K).astype(np.float32))
    b = _bf(rs.randn(K, N).astype(np.float32) * 0.1)
    # leading dimensions padded to multiples of 8 elements (TMA pitch must be a multiple of 16 bytes)
    def put(x):
        r, c = x.shape
        buf = torch.zeros(r, (c + 7) // 8 * 8, dtype=torch.bfloat16, device="cuda")
        buf[:, :c] = torch.tensor(x, dtype=torch.bfloat16)
        return buf[:, :c] if False else buf
    A = put(a.T.copy() if ta else a)
    B = put(b.T.copy() if tb else b)
    ref = a.astype(np.float64) @ b.astype(np.float64)
    got = _gemm(lib, ctx, 1, ta, tb, A, B, M, N, K).cpu().numpy()
    assert relerr(got, ref) < 2e-5, "fp32-accumulated bf16 GEMM must match fp64 on bf16-exact inputs"
    if K >= 256:
        Cm = torch.zeros(M, N, device="cuda")
        got = _gemm(lib, ctx, 1, ta, tb, A, B, M, N, K, accumulate=1, split_k=3, C=Cm).cpu().numpy()
        assert relerr(got, ref) < 2e-5


def test_gemm_bf16_epilogue(lib, ctx):
    rs = np.random.RandomState(3)
    M, N, K = 200, 128, 192
    a, b = _bf(rs.randn(M, K).astype(np.float32)), _bf(rs.randn(K, N).astype(np.float32) * 0.1)
    mask = _bf(rs.randn(M, N).astype(np.float32))
    A, B, Mk = dev(a, torch.bfloat16), dev(b, torch.bfloat16), dev(mask, torch.bfloat16)
    ref = a.astype(np.float64) @ b
    r1 = np.maximum(ref, 0)
    j = np.arange(N) % 64
    r1[:, j == 50] = 1.0
    r1[:, j > 50] = 0.0
    got = _gemm(lib, ctx, 1, 0, 0, A, B, M, N, K, out_dtype=1, act=1, n_valid=50, n_block=64).float().cpu().numpy()
    assert relerr(got, r1) < 1e-2          # bf16 output rounding
    r2 = ref * (mask > 0)
    r2[:, j >= 50] = 0.0
    got = _gemm(lib, ctx, 1, 0, 0, A, B, M, N, K, out_dtype=0, mask=Mk, n_valid=50, n_block=64, pad_one=0.0).cpu().numpy()
    assert relerr(got, r2) < 2e-5


# ---------------------------------------------------------------------------------------------
# fused ELBO
# ---------------------------------------------------------------------------------------------
def _run_elbo(lib, ctx, mode, input_type, X, dec, mean, lv, logits, eps, zeta, tau, m, plv, r, s, x_dtype=0, dec_dtype=0,
              recon_scale=1.0, x_scale=0.0):
    from dmvae_b200 import _abi
    B, D = X.shape
    L, K = mean.shape[1], m.shape[0]
    tdt = {0: torch.float32, 1: torch.bfloat16, 2: torch.uint8}
    Dp = (D + 1 + 63) // 64 * 64
    Xd = torch.zeros(B, (D + 15) // 16 * 16, dtype=tdt[x_dtype], device="cuda")
    Xd[:, :D] = torch.tensor(X).to(tdt[x_dtype])
    dd = torch.zeros(B, Dp, dtype=tdt[dec_dtype], device="cuda")
    dd[:, :D] = torch.tensor(dec).to(tdt[dec_dtype])
    zh = torch.zeros(B, 2 * L, device="cuda")
    zh[:, :L] = dev(mean)
    zh[:, L:] = dev(lv)
    ea = _abi.ElboArgs()
    ea.mode, ea.input_type, ea.rows, ea.D, ea.L, ea.K = mode, input_type, B, D, L, K
    ea.X, ea.x_dtype, ea.ldx = Xd.data_ptr(), x_dtype, Xd.stride(0)
    ea.decoded, ea.dec_dtype, ea.ld_dec = dd.data_ptr(), dec_dtype, dd.stride(0)
    ea.mean, ea.log_var, ea.ld_zh = zh.data_ptr(), zh.data_ptr() + 4 * L, 2 * L
    keep = [Xd, dd, zh]
    out = {}
    if logits is not None:
        lg = dev(logits); keep.append(lg)
        ea.logits, ea.ld_logits = lg.data_ptr(), K
        Kp = (K + 63) // 64 * 64
        out["d_logits"] = torch.full((B, Kp), 7.0, dtype=tdt[dec_dtype], device="cuda")
        ea.d_logits, ea.dlogits_dtype, ea.ld_dlogits, ea.dlogits_cols = out["d_logits"].data_ptr(), dec_dtype, Kp, Kp
    if eps is not None:
        ep = dev(eps); keep.append(ep)
        ea.eps, ea.ld_eps = ep.data_ptr(), L
    if zeta is not None:
        zt = dev(zeta); keep.append(zt)
        ea.zeta, ea.ld_zeta = zt.data_ptr(), K
    ea.tau = tau
    pm, pl = dev(m), dev(plv)
    ea.prior_means, ea.prior_log_vars = pm.data_ptr(), pl.data_ptr()
    ea.kl_ratio, ea.inv_global_batch, ea.recon_scale = r, s, recon_scale
    ea.x_scale = x_scale
    out["per_sample"] = torch.zeros(B, 4, device="cuda")
    out["qc"] = torch.zeros(B, K, device="cuda")
    out["argmax"] = torch.zeros(B, dtype=torch.int32, device="cuda")
    out["d_decoded"] = torch.full((B, Dp), 7.0, dtype=tdt[dec_dtype], device="cuda")
    out["d_mean"] = torch.zeros(B, L, device="cuda")
    out["d_log_var"] = torch.zeros(B, L, device="cuda")
    out["d_Z_gamma"] = torch.zeros(B, L, device="cuda")
    w_s, f_s = torch.zeros(B, K, device="cuda"), torch.zeros(B, 2 * L, device="cuda")
    ea.per_sample, ea.qc, ea.argmax = out["per_sample"].data_ptr(), out["qc"].data_ptr(), out["argmax"].data_ptr()
    ea.d_decoded, ea.ld_ddec, ea.ddec_cols = out["d_decoded"].data_ptr(), Dp, Dp
    ea.d_mean_kl, ea.d_log_var_kl, ea.ld_dkl = out["d_mean"].data_ptr(), out["d_log_var"].data_ptr(), L
    ea.d_Z_gamma, ea.ld_dzg = out["d_Z_gamma"].data_ptr(), L
    ea.w_scratch, ea.f_scratch = w_s.data_ptr(), f_s.data_ptr()
    _abi.check(lib.dmvae_elbo_fwd_bwd(ctx, C.byref(ea), stream()))
    out["d_means"] = torch.zeros(K, L, device="cuda")
    out["d_log_vars"] = torch.zeros(K, L, device="cuda")
    out["loss"] = torch.zeros(4, device="cuda")
    ws = torch.zeros(int(lib.dmvae_elbo_reduce_workspace(B, L, K)), device="cuda")
    _abi.check(lib.dmvae_elbo_reduce(ctx, C.byref(ea), out["d_means"].data_ptr(), out["d_log_vars"].data_ptr(), 0,
                                     out["loss"].data_ptr(), ws.data_ptr(), stream()))
    torch.cuda.synchronize()
    return {k: v.float().cpu().numpy() if v.dtype != torch.int32 else v.cpu().numpy() for k, v in out.items()}


def _elbo_inputs(B, D, L, K, seed, binary=True):
    rs = np.random.RandomState(seed)
    X = (rs.uniform(size=(B, D)) < 0.3).astype(np.float32) if binary else rs.uniform(size=(B, D)).astype(np.float32)
    dec = (rs.randn(B, D) * 2).astype(np.float32)
    mean, lv = rs.randn(B, L).astype(np.float32), (rs.randn(B, L) * 0.5).astype(np.float32)
    logits = (rs.randn(B, K) * 2).astype(np.float32)
    eps = rs.randn(B, L).astype(np.float32)
    m, plv = rs.randn(K, L).astype(np.float32), (rs.randn(K, L) * 0.3).astype(np.float32)
    return X, dec, mean, lv, logits, eps, m, plv


def _check_common(got, c, D, tol=1e-4, ps_abs=2e-5):
    ref_ps = np.stack([c["R"], c["C"], c["Zk"], c["elbo"]], 1)
    assert np.all(np.abs(got["per_sample"] - ref_ps) <= tol * np.abs(ref_ps) + ps_abs), \
        "per-sample ELBO terms beyond 1e-4 relative (+2e-5 absolute fp32 rounding floor)"
    assert relerr(got["qc"], c["q"]) < tol
    assert np.array_equal(got["argmax"], c["argmax"]), "cluster assignments must be bit-exact"
    assert relerr(got["d_decoded"][:, :D], c["d_decoded"]) < tol
    assert np.all(got["d_decoded"][:, D:] == 0)
    assert relerr(got["d_mean"], c["d_mean"]) < tol
    assert relerr(got["d_log_var"], c["d_log_var"]) < tol
    assert relerr(got["d_means"], c["d_means"]) < tol
    assert relerr(got["d_log_vars"], c["d_log_vars"]) < tol
    assert abs(got["loss"][3] - c["loss"]) < tol * abs(c["loss"])


@pytest.mark.parametrize("B,D,L,K", [(256, 784, 10, 10), (37, 12, 3, 4), (130, 3072, 128, 100), (64, 784, 64, 50),
                                     (100, 784, 10, 10), (45, 64, 5, 7), (1, 784, 10, 10), (4096, 784, 10, 10)])
@pytest.mark.parametrize("binary", [True, False])
def test_elbo_dmvae_fp32(lib, ctx, B, D, L, K, binary):
    X, dec, mean, lv, logits, eps, m, plv = _elbo_inputs(B, D, L, K, 0, binary)
    it = 0 if binary else 1
    r, s = 0.7, 1.0 / (2 * B)
    got = _run_elbo(lib, ctx, 0, it, X, dec, mean, lv, logits, None, None, 1.0, m, plv, r, s)
    c = cf.elbo_dmvae(*[a.astype(np.float64) for a in (X, dec, mean, lv, logits, m, plv)], r=r, s=s,
                      input_type="binary" if binary else "real")
    _check_common(got, c, D)
    assert relerr(got["d_logits"][:, :K], c["d_logits"]) < 1e-4
    assert np.all(got["d_logits"][:, K:] == 0)


@pytest.mark.parametrize("B", [256, 17, 4095, 1])
def test_elbo_u8_and_bf16_variants(lib, ctx, B):
    """uint8 targets and the bf16 tier of the row-tile kernel (the benchmarked variant), including ragged row counts:
    17 = one full 16-row tile + a 1-row tile, 4095 = the last tile one row short."""
    D, L, K = 784, 10, 10
    X, dec, mean, lv, logits, eps, m, plv = _elbo_inputs(B, D, L, K, 1, True)
    c = cf.elbo_dmvae(*[a.astype(np.float64) for a in (X, dec, mean, lv, logits, m, plv)])
    got = _run_elbo(lib, ctx, 0, 0, X, dec, mean, lv, logits, None, None, 1.0, m, plv, 1.0, 1.0 / B, x_dtype=2)
    _check_common(got, c, D)                        # uint8 targets are exact for binarised data
    decb = _bf(dec)
    cb = cf.elbo_dmvae(*[a.astype(np.float64) for a in (X, decb, mean, lv, logits, m, plv)])
    got = _run_elbo(lib, ctx, 0, 0, X, dec, mean, lv, logits, None, None, 1.0, m, plv, 1.0, 1.0 / B, x_dtype=2, dec_dtype=1)
    ref_ps = np.stack([cb["R"], cb["C"], cb["Zk"], cb["elbo"]], 1)
    # bf16 tier (north_star tolerance 2e-2): the reconstruction term uses the one-MUFU tanh formulation in fp32;
    # measured error is ~1e-5, asserted at 1e-3.  The latent terms stay fp32 (1e-4).
    assert np.array_equal(got["argmax"], cb["argmax"])
    assert np.all(got["d_decoded"][:, D:] == 0) and np.all(got["d_logits"][:, K:] == 0)
    assert np.all(np.abs(got["per_sample"][:, 1:3] - ref_ps[:, 1:3]) <= 1e-4 * np.abs(ref_ps[:, 1:3]) + 2e-5)
    assert np.all(np.abs(got["per_sample"] - ref_ps) <= 1e-3 * np.abs(ref_ps) + 2e-5)
    assert relerr(got["d_decoded"][:, :D], cb["d_decoded"]) < 1e-2          # bf16 store
    assert relerr(got["d_logits"][:, :K], cb["d_logits"]) < 1e-2


@pytest.mark.parametrize("B,D,L,K", [(100, 784, 10, 10), (70, 3072, 40, 24)])
@pytest.mark.parametrize("dec_dtype", [0, 1])
def test_elbo_uint8_intensities_with_scale(lib, ctx, B, D, L, K, dec_dtype):
    """8-bit targets x = byte / 255 (the CIFAR-shaped soft targets of includes/utils.py:204-210) read as uint8 with
    x_scale = 1/255: row-tile kernel (K*L <= 512) and the MMA + streaming kernels, fp32 and bf16 decoder logits."""
    X, dec, mean, lv, logits, eps, m, plv = _elbo_inputs(B, D, L, K, 7, True)
    Xb = np.random.RandomState(8).randint(0, 256, size=(B, D)).astype(np.float32)         # byte values
    dec_ref = _bf(dec) if dec_dtype == 1 else dec
    c = cf.elbo_dmvae(*[a.astype(np.float64) for a in (Xb / 255.0, dec_ref, mean, lv, logits, m, plv)], r=0.9, s=1.0 / B)
    got = _run_elbo(lib, ctx, 0, 0, Xb, dec, mean, lv, logits, None, None, 1.0, m, plv, 0.9, 1.0 / B, x_dtype=2,
                    dec_dtype=dec_dtype, x_scale=1.0 / 255.0)
    ref_ps = np.stack([c["R"], c["C"], c["Zk"], c["elbo"]], 1)
    tol = 1e-4 if dec_dtype == 0 else 1e-3
    assert np.all(np.abs(got["per_sample"] - ref_ps) <= tol * np.abs(ref_ps) + 2e-5)
    assert relerr(got["d_decoded"][:, :D], c["d_decoded"]) < (1e-4 if dec_dtype == 0 else 1e-2)
    assert np.array_equal(got["argmax"], c["argmax"])


def test_elbo_sampled(lib, ctx):
    B, D, L, K = 200, 784, 10, 10
    X, dec, mean, lv, logits, eps, m, plv = _elbo_inputs(B, D, L, K, 2, True)
    rs = np.random.RandomState(5)
    gum = rg.sample_gumbel(rs, (B, K)); tau = 0.6
    c = cf.elbo_dmvae_sampled(*[a.astype(np.float64) for a in (X, dec, mean, lv, logits)], gum, tau,
                              m.astype(np.float64), plv.astype(np.float64), r=0.9)
    got = _run_elbo(lib, ctx, 1, 0, X, dec, mean, lv, logits, None, c["zeta"], tau, m, plv, 0.9, 1.0 / B)
    _check_common(got, c, D)
    assert relerr(got["d_logits"][:, :K], c["d_logits"]) < 1e-4


@pytest.mark.parametrize("B,D,L,K", [(256, 784, 64, 50), (33, 12, 3, 4)])
def test_elbo_vade(lib, ctx, B, D, L, K):
    X, dec, mean, lv, logits, eps, m, plv = _elbo_inputs(B, D, L, K, 3, True)
    mean *= 0.5; m *= 0.5
    Z = mean.astype(np.float64) + np.exp(lv.astype(np.float64) / 2) * eps
    c = cf.elbo_vade(X.astype(np.float64), dec.astype(np.float64), mean.astype(np.float64), lv.astype(np.float64), Z,
                     m.astype(np.float64), plv.astype(np.float64), r=0.8)
    got = _run_elbo(lib, ctx, 2, 0, X, dec, mean, lv, None, eps, None, 1.0, m, plv, 0.8, 1.0 / B)
    # gamma is a softmax of O(L)-sized scores: allow the fp32 rounding of the scores in q
    _check_common(got, c, D, tol=3e-4)
    assert relerr(got["d_Z_gamma"], c["d_Z_gamma"]) < 3e-4


def test_elbo_rejects_bad_input_type(lib, ctx):
    from dmvae_b200 import _abi
    ea = _abi.ElboArgs()
    ea.input_type = 5
    ea.mode, ea.rows, ea.D, ea.L, ea.K = 0, 4, 8, 2, 2
    assert lib.dmvae_elbo_fwd_bwd(ctx, C.byref(ea), stream()) == 1            # DMVAE_ERR_INVALID
    assert b"not implemented" in lib.dmvae_last_error()                       # base_models.py:84-85
    ea.input_type = 0
    assert lib.dmvae_elbo_fwd_bwd(ctx, C.byref(ea), stream()) == 1            # NULL inputs are rejected, not dereferenced
    assert b"NULL" in lib.dmvae_last_error()


# ---------------------------------------------------------------------------------------------
# reparameterisation / Philox
# ---------------------------------------------------------------------------------------------
def _reparam(lib, ctx, mean, lv, logits, eps_in, gum_in, seed, step, row_offset, tau, z_dtype=0):
    from dmvae_b200 import _abi
    B, L = mean.shape
    K = logits.shape[1] if logits is not None else 0
    zh = torch.zeros(B, 2 * L, device="cuda"); zh[:, :L] = dev(mean); zh[:, L:] = dev(lv)
    ra = _abi.ReparamArgs()
    ra.rows, ra.L, ra.K = B, L, K
    ra.mean, ra.log_var, ra.ld_zh = zh.data_ptr(), zh.data_ptr() + 4 * L, 2 * L
    keep = []
    if logits is not None:
        lg = dev(logits); keep.append(lg); ra.logits, ra.ld_logits = lg.data_ptr(), K
    if eps_in is not None:
        e = dev(eps_in); keep.append(e); ra.eps_in = e.data_ptr()
    if gum_in is not None:
        g = dev(gum_in); keep.append(g); ra.gumbel_in = g.data_ptr()
    ra.seed, ra.step, ra.row_offset, ra.tau = seed, step, row_offset, tau
    Z = torch.full((B, 64), 9.0, dtype=torch.float32 if z_dtype == 0 else torch.bfloat16, device="cuda")
    eps_out = torch.zeros(B, L, device="cuda")
    zeta = torch.zeros(B, max(K, 1), device="cuda")
    ra.Z_out, ra.z_dtype, ra.ld_z, ra.z_cols = Z.data_ptr(), z_dtype, 64, 64
    ra.eps_out = eps_out.data_ptr()
    ra.zeta_out = zeta.data_ptr() if logits is not None else None
    _abi.check(lib.dmvae_reparam_fwd(ctx, C.byref(ra), stream()))
    torch.cuda.synchronize()
    return Z.float().cpu().numpy(), eps_out.cpu().numpy(), zeta.cpu().numpy()


def test_reparam_injected_and_padding(lib, ctx):
    rs = np.random.RandomState(0)
    B, L, K = 300, 10, 10
    mean, lv, logits = rs.randn(B, L).astype(np.float32), rs.randn(B, L).astype(np.float32) * .4, rs.randn(B, K).astype(np.float32)
    eps, gum = rs.randn(B, L).astype(np.float32), rg.sample_gumbel(rs, (B, K)).astype(np.float32)
    Z, eo, zeta = _reparam(lib, ctx, mean, lv, logits, eps, gum, 0, 0, 0, 0.5)
    assert relerr(Z[:, :L], mean + np.exp(lv.astype(np.float64) / 2) * eps) < 1e-6
    assert np.all(Z[:, L] == 1.0) and np.all(Z[:, L + 1:] == 0.0)
    assert np.array_equal(eo, eps)
    assert relerr(zeta, cf.softmax((logits.astype(np.float64) + gum) / 0.5)) < 1e-5
    Z0, _, _ = _reparam(lib, ctx, mean, lv, None, np.zeros_like(eps), None, 0, 0, 0, 1.0)
    assert np.array_equal(Z0[:, :L], mean)                               # eps = 0 -> Z = mu (pre-training)


def test_reparam_philox_matches_oracle(lib, ctx):
    rs = np.random.RandomState(1)
    B, L, K = 512, 10, 10
    mean, lv, logits = np.zeros((B, L), np.float32), np.zeros((B, L), np.float32), rs.randn(B, K).astype(np.float32)
    seed, step, off = 0x1234567887654321, 17, 1000
    Z, eo, zeta = _reparam(lib, ctx, mean, lv, logits, None, None, seed, step, off, 1.0)
    ref = philox.normal(B, L, seed, step, off)
    assert np.abs(eo - ref).max() < 2e-5            # same integer stream; fp32 log/sincos vs fp64
    g = philox.gumbel(B, K, seed, step, off)
    assert relerr(zeta, cf.softmax(logits.astype(np.float64) + g)) < 2e-4
    # L not a multiple of 4, K > 32
    Z, eo, zeta = _reparam(lib, ctx, np.zeros((64, 7), np.float32), np.zeros((64, 7), np.float32),
                           np.zeros((64, 50), np.float32), None, None, 5, 3, 0, 1.0)
    assert np.abs(eo - philox.normal(64, 7, 5, 3)).max() < 2e-5
    assert relerr(zeta, cf.softmax(philox.gumbel(64, 50, 5, 3))) < 2e-4


def test_reparam_bwd(lib, ctx):
    from dmvae_b200 import _abi
    rs = np.random.RandomState(2)
    B, L = 100, 10
    a = [rs.randn(B, L).astype(np.float32) for _ in range(6)]
    dmk, dlk, dZ, dZe, eps, lv = a
    dz = torch.zeros(B, 64, device="cuda"); dz[:, :L] = dev(dZ)
    zh = torch.zeros(B, 2 * L, device="cuda"); zh[:, L:] = dev(lv)
    out = torch.full((B, 64), 5.0, device="cuda")
    t = [dev(x) for x in (dmk, dlk, dZe, eps)]
    _abi.check(lib.dmvae_reparam_bwd(ctx, B, L, t[0].data_ptr(), t[1].data_ptr(), L, dz.data_ptr(), 64, t[2].data_ptr(), L,
                                     t[3].data_ptr(), zh.data_ptr() + 4 * L, 2 * L, None, 0, out.data_ptr(), 0, 64, 64, stream()))
    torch.cuda.synchronize()
    o = out.cpu().numpy()
    rm, rl = cf.reparam_backward(dmk.astype(np.float64), dlk.astype(np.float64), dZ.astype(np.float64) + dZe, eps, lv.astype(np.float64))
    assert relerr(o[:, :L], rm) < 1e-6 and relerr(o[:, L:2 * L], rl) < 1e-5 and np.all(o[:, 2 * L:] == 0)


# ---------------------------------------------------------------------------------------------
# Adam / staging / eval
# ---------------------------------------------------------------------------------------------
def test_adam_tf_semantics(lib, ctx):
    from dmvae_b200 import _abi
    rs = np.random.RandomState(3)
    n = 4096 + 8
    th, g = rs.randn(n).astype(np.float32), rs.randn(n).astype(np.float32)
    g[::7] = 0.0
    p, gr = dev(th), dev(g)
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    pb = torch.zeros(n, dtype=torch.bfloat16, device="cuda")
    th64, m64, v64 = th.astype(np.float64), np.zeros(n), np.zeros(n)
    for t in range(1, 4):
        lr_t = 0.002 * math.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)
        gr.copy_(dev(g))
        _abi.check(lib.dmvae_adam(ctx, p.data_ptr(), gr.data_ptr(), m.data_ptr(), v.data_ptr(), pb.data_ptr(), n, lr_t, None, 0.9,
                                  0.999, 1e-8, 1.0, 1, stream()))
        rg.adam_tf_step(th64, g.astype(np.float64), m64, v64, t, 0.002)
    torch.cuda.synchronize()
    assert relerr(p.cpu().numpy(), th64) < 1e-6
    assert np.all(gr.cpu().numpy() == 0)
    assert np.array_equal(pb.float().cpu().numpy(), p.to(torch.bfloat16).float().cpu().numpy())
    assert np.all(p.cpu().numpy()[::7] == th[::7])          # zero gradient from the start -> never moves


def test_adam_launch_shapes_and_exchange_kernel_agree(lib, ctx):
    """The background launch shape of Adam (4-warp blocks that fit beside a GEMM CTA) and the data-parallel exchange
    kernel with world = 1 (both launch shapes) are the same arithmetic as dmvae_adam: bit-identical parameters, slots,
    bf16 operand copy; a sub-range call (pointer offsets) only touches its range."""
    from dmvae_b200 import _abi
    rs = np.random.RandomState(11)
    n = 64 * 64 * 3 + 64
    th, g = rs.randn(n).astype(np.float32), (rs.randn(n) * 1e-2).astype(np.float32)
    m0, v0 = (rs.randn(n) * 1e-3).astype(np.float32), (rs.rand(n) * 1e-4).astype(np.float32)

    def run(kind):
        p, gr, m, v = dev(th), dev(g), dev(m0), dev(v0)
        pb = torch.zeros(n, dtype=torch.bfloat16, device="cuda")
        if kind in ("fg", "bg"):
            _abi.check(lib.dmvae_adam(ctx, p.data_ptr(), gr.data_ptr(), m.data_ptr(), v.data_ptr(), pb.data_ptr(), n, 1.5e-3,
                                      None, 0.9, 0.999, 1e-8, 1.0, 1 if kind == "fg" else 3, stream()))
        elif kind == "range":                    # three calls over disjoint sub-ranges, mixed shapes
            for lo, hi, fl in ((0, 4096, 3), (4096, 8192 + 64, 1), (8192 + 64, n, 3)):
                _abi.check(lib.dmvae_adam(ctx, p.data_ptr() + 4 * lo, gr.data_ptr() + 4 * lo, m.data_ptr() + 4 * lo,
                                          v.data_ptr() + 4 * lo, pb.data_ptr() + 2 * lo, hi - lo, 1.5e-3, None, 0.9, 0.999,
                                          1e-8, 1.0, fl, stream()))
        else:                                    # exchange kernel, one rank: its own buffers are the only "peers"
            VP = C.c_void_p * 1
            _abi.check(lib.dmvae_dp_reduce_adam(ctx, 0, 1, VP(gr.data_ptr()), VP(p.data_ptr()), VP(pb.data_ptr()),
                                                None, None, 0, m.data_ptr(), v.data_ptr(), n, 0, n, 1.5e-3, None, 0.9, 0.999, 1e-8,
                                                1 | (2 if kind == "dp_bg" else 0), stream()))
        torch.cuda.synchronize()
        return [t.clone() for t in (p, m, v, pb, gr)]

    ref = run("fg")
    assert float(ref[4].abs().max()) == 0.0
    for kind in ("bg", "range", "dp", "dp_bg"):
        got = run(kind)
        for a, b in zip(ref, got):
            assert torch.equal(a, b), kind
    assert lib.dmvae_adam(ctx, ref[0].data_ptr(), ref[4].data_ptr(), ref[1].data_ptr(), ref[2].data_ptr(), None, n, 1e-3, None,
                          0.9, 0.999, 1e-8, 1.0, 8, stream()) != 0          # unknown flag bits are rejected


def test_step_tick_and_device_lr(lib, ctx):
    """CUDA-graph replay reads Adam's lr_t and the Philox step from device memory."""
    from dmvae_b200 import _abi
    st = np.zeros(1, dtype=[("step", "<u8"), ("t", "<u4"), ("lr_t", "<f4")])
    st["step"], st["t"] = 41, 6
    sd = torch.from_numpy(st.view(np.int32).copy()).cuda()
    _abi.check(lib.dmvae_step_tick(ctx, sd.data_ptr(), 0.002, 0.9, 0.999, stream()))
    torch.cuda.synchronize()
    back = sd.cpu().numpy().view(st.dtype)
    assert back["step"][0] == 42 and back["t"][0] == 7
    assert abs(back["lr_t"][0] - 0.002 * math.sqrt(1 - 0.999 ** 7) / (1 - 0.9 ** 7)) < 1e-8
    n = 1024
    p, g = torch.ones(n, device="cuda"), torch.ones(n, device="cuda")
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    _abi.check(lib.dmvae_adam(ctx, p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), None, n, 123.0,
                              sd.data_ptr() + 12, 0.9, 0.999, 1e-8, 1.0, 0, stream()))
    torch.cuda.synchronize()
    exp = 1.0 - float(back["lr_t"][0]) * 0.1 / (math.sqrt(0.001) + 1e-8)
    assert abs(float(p[0]) - exp) < 1e-6


def test_stage_input_and_argmax(lib, ctx):
    from dmvae_b200 import _abi
    rs = np.random.RandomState(4)
    B, D = 77, 784
    X = (rs.uniform(size=(B, D)) < .13).astype(np.uint8)
    Xd = torch.tensor(X, device="cuda")
    A = torch.full((B, 832), 3.0, dtype=torch.bfloat16, device="cuda")
    _abi.check(lib.dmvae_stage_input(ctx, Xd.data_ptr(), 2, D, A.data_ptr(), 1, 832, B, D, 1.0, stream()))
    a = A.float().cpu().numpy()
    assert np.array_equal(a[:, :D], X) and np.all(a[:, D] == 1) and np.all(a[:, D + 1:] == 0)
    sc = rs.randn(1000, 10).astype(np.float32); sc[5, 3] = sc[5, 7] = 99.0
    cls = rs.randint(0, 10, 1000).astype(np.int32)
    am = torch.zeros(1000, dtype=torch.int32, device="cuda"); cnt = torch.zeros(10, 10, dtype=torch.int32, device="cuda")
    s_, c_ = dev(sc), torch.tensor(cls, device="cuda")
    _abi.check(lib.dmvae_argmax_contingency(ctx, s_.data_ptr(), 10, 1000, 10, c_.data_ptr(), 10, am.data_ptr(), cnt.data_ptr(), stream()))
    torch.cuda.synchronize()
    ref = np.argmax(sc, 1)
    assert np.array_equal(am.cpu().numpy(), ref)
    d = np.zeros((10, 10), np.int32)
    for i in range(1000):
        d[ref[i], cls[i]] += 1
    assert np.array_equal(cnt.cpu().numpy(), d)
