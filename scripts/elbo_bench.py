#!/usr/bin/env python
"""Time the fused ELBO fwd+bwd kernel alone (u8 targets, bf16 decoder logits / gradient: the bench.py variant).

    python scripts/elbo_bench.py [--rows 4096,65536] [--D 784] [--L 10] [--K 10] [--iters 20]

Prints device time per launch (CUDA-graph replay of `iters` launches over rotating buffers that exceed L2 for the
large sizes), algorithmic GB/s (SURVEY 8d byte count) and the fraction of MEASURED_PEAKS.json's copy bandwidth.
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dmvae_b200 import _abi  # noqa: E402


def time_elbo(lib, ctx, B, D=784, L=10, K=10, iters=20, check=False, log=print):
    """Device time (us) of one fused-ELBO launch at B rows: CUDA-graph replay of `iters` launches over rotating buffer
    sets (> L2 when the size allows)."""
    Dp = (D + 1 + 63) // 64 * 64
    Kp = (K + 63) // 64 * 64
    nbuf = max(2, min(8, int(400e6 // (B * D * 5)) + 1))        # rotate over > 126 MB when the size allows
    sets = []
    for i in range(nbuf):
        X = (torch.rand(B, D, device="cuda") < 0.13).to(torch.uint8)
        dec = (torch.randn(B, Dp, device="cuda") * 2).to(torch.bfloat16)
        ddec = torch.empty_like(dec)
        sets.append((X, dec, ddec))
    Zp = (2 * L + 63) // 64 * 64
    zh = torch.randn(B, Zp, device="cuda") * 0.5
    lg = torch.randn(B, Kp, device="cuda")
    dlg = torch.empty(B, Kp, dtype=torch.bfloat16, device="cuda")
    pm, pl = torch.randn(K, L, device="cuda"), torch.randn(K, L, device="cuda") * 0.3
    ps, qc = torch.empty(B, 4, device="cuda"), torch.empty(B, K, device="cuda")
    am = torch.empty(B, dtype=torch.int32, device="cuda")
    dm, dl = torch.empty(B, L, device="cuda"), torch.empty(B, L, device="cuda")
    eas = []
    for X, dec, ddec in sets:
        ea = _abi.ElboArgs()
        ea.mode, ea.input_type, ea.rows, ea.D, ea.L, ea.K = 0, 0, B, D, L, K
        ea.X, ea.x_dtype, ea.ldx = X.data_ptr(), 2, D
        ea.decoded, ea.dec_dtype, ea.ld_dec = dec.data_ptr(), 1, Dp
        ea.mean, ea.log_var, ea.ld_zh = zh.data_ptr(), zh.data_ptr() + 4 * L, Zp
        ea.logits, ea.ld_logits = lg.data_ptr(), Kp
        ea.d_logits, ea.dlogits_dtype, ea.ld_dlogits, ea.dlogits_cols = dlg.data_ptr(), 1, Kp, Kp
        ea.prior_means, ea.prior_log_vars = pm.data_ptr(), pl.data_ptr()
        ea.kl_ratio, ea.inv_global_batch, ea.recon_scale, ea.tau = 1.0, 1.0 / B, 1.0, 1.0
        ea.per_sample, ea.qc, ea.argmax = ps.data_ptr(), qc.data_ptr(), am.data_ptr()
        ea.d_decoded, ea.ld_ddec, ea.ddec_cols = ddec.data_ptr(), Dp, Dp
        ea.d_mean_kl, ea.d_log_var_kl, ea.ld_dkl = dm.data_ptr(), dl.data_ptr(), L
        eas.append(ea)
    st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for ea in eas:
        _abi.check(lib.dmvae_elbo_fwd_bwd(ctx, C.byref(ea), st()))
    torch.cuda.synchronize()
    if check:
        X, dec, ddec = sets[0]
        for name in ("randn*2", "trained-like"):
            if name == "trained-like":                       # confident logits that mostly agree with the targets
                dec[:, :D] = ((X.float() * 2 - 1) * (6 + 3 * torch.randn(B, D, device="cuda"))).to(torch.bfloat16)
            _abi.check(lib.dmvae_elbo_fwd_bwd(ctx, C.byref(eas[0]), st()))
            torch.cuda.synchronize()
            d64, x64 = dec[:, :D].double(), X.double()
            R = (torch.nn.functional.softplus(d64) - d64 * x64).sum(1)
            g = (torch.sigmoid(d64) - x64) / B
            eR = ((ps[:, 0].double() - R).abs() / R.abs()).max().item()
            eg = (ddec[:, :D].double() - g).abs().max().item() * B
            log("  check %-12s rows %d: recon term max rel err %.2e (mean R %.1f), d_decoded max abs err %.2e x 1/B"
                % (name, B, eR, R.mean().item(), eg))
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            _abi.check(lib.dmvae_elbo_fwd_bwd(ctx, C.byref(eas[i % nbuf]), st()))
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    # the cross-sample reduction of the same pass (prior-table gradients + loss terms), timed the same way
    gm, gl, lo = torch.zeros(K, L, device="cuda"), torch.zeros(K, L, device="cuda"), torch.zeros(4, device="cuda")
    ws = torch.zeros(int(lib.dmvae_elbo_reduce_workspace(B, L, K)), device="cuda")
    red = lambda: _abi.check(lib.dmvae_elbo_reduce(ctx, C.byref(eas[0]), gm.data_ptr(), gl.data_ptr(), 0, lo.data_ptr(), ws.data_ptr(), st()))
    red()
    torch.cuda.synchronize()
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g2):
        for i in range(iters):
            red()
    g2.replay()
    torch.cuda.synchronize()
    e0.record()
    g2.replay()
    e1.record()
    torch.cuda.synchronize()
    log("  elbo_reduce (2 kernels) rows %d L %d K %d: %.2f us" % (B, L, K, e0.elapsed_time(e1) * 1e3 / iters))
    return us, nbuf


def elbo_bytes_per_sample(D, L, K):
    """u8 targets, bf16 decoder logits in, bf16 gradient out, fp32 latent I/O."""
    return 5 * D + 4 * (3 * L + K) + 4 * (2 * L + K) + 4 * K + 4 * L + 12


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", default="4096,65536")
    ap.add_argument("--D", type=int, default=784)
    ap.add_argument("--L", type=int, default=10)
    ap.add_argument("--K", type=int, default=10)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--check", action="store_true", help="report the error of the reconstruction term and of d_decoded against "
                    "torch fp64 on the same bf16 logits (randn logits and trained-like saturated logits)")
    args = ap.parse_args()
    peak = 6545.9
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk))["hbm_gbs"]
    lib = _abi.load()
    ctx = C.c_void_p()
    _abi.check(lib.dmvae_ctx_create(0, C.byref(ctx)))
    D, L, K = args.D, args.L, args.K
    for B in [int(x) for x in args.rows.split(",")]:
        us, nbuf = time_elbo(lib, ctx, B, D, L, K, args.iters, args.check)
        bytes_ps = elbo_bytes_per_sample(D, L, K)
        gbs = B * bytes_ps / us * 1e-3
        print("rows %6d D %d L %d K %d: %8.2f us/launch, %7.1f GB/s algorithmic (%d B/sample), %.3f of measured %.1f GB/s; "
              "%d rotating buffer sets" % (B, D, L, K, us, gbs, bytes_ps, gbs / peak, peak, nbuf))
    lib.dmvae_ctx_destroy(ctx)


if __name__ == "__main__":
    main()
