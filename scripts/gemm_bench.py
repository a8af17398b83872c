#!/usr/bin/env python
"""Validate and time the bf16 tcgen05 GEMM (dmvae_gemm) on every shape the DMVAE step launches (cfg2: batch 4096).

    python scripts/gemm_bench.py [--rows 4096] [--iters 30] [--check]

For each (layer, pass) it prints the device time per launch (CUDA events around `iters` back-to-back launches),
TFLOP/s on the padded shape, and with --check the max error against a torch fp32 matmul of the same bf16 operands.
DMVAE_GEMM_PAIR=0 in the environment selects the single-CTA kernel for comparison.
"""
import argparse
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dmvae_b200 import _abi  # noqa: E402
from dmvae_b200._abi import BF16, F32  # noqa: E402

# (name, n_in_pad, n_out_pad) of the dense layers at cfg2 (engine.Layout)
LAYERS = [("enc1", 832, 512), ("enc2", 512, 512), ("ench", 512, 4096), ("zh", 2048, 64), ("ch", 2048, 64),
          ("dec1", 64, 2048), ("dec2", 2048, 512), ("dec3", 512, 512), ("decx", 512, 832)]


def time_gemms(lib, ctx, rows=4096, iters=30, check=False, only="", dev=None, verbose=True):
    """Device time of every GEMM of one training step (27 launches at cfg2), each as a CUDA-graph replay of `iters`
    back-to-back launches.  Returns (sum of us per step, padded FLOP per step, worst relative error or nan)."""
    class _A:
        pass
    args = _A()
    args.rows, args.iters, args.check, args.only = rows, iters, check, only
    dev = torch.device("cuda", torch.cuda.current_device()) if dev is None else dev
    torch.manual_seed(0)
    B = args.rows
    _print = print if verbose else (lambda *a, **k: None)
    st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
    total_us, total_flop, worst = 0.0, 0.0, 0.0
    _print("%-12s %-6s %6s %6s %6s %5s %9s %9s %10s" % ("layer", "pass", "M", "N", "K", "split", "us", "TFLOP/s", "max_err"))
    for name, kin, nout in LAYERS:
        if args.only and name not in args.only.split(","):
            continue
        X = (torch.randn(B, kin, device=dev) * 0.5).to(torch.bfloat16)
        W = (torch.randn(kin, nout, device=dev) * 0.05).to(torch.bfloat16)
        dY = (torch.randn(B, nout, device=dev) * 0.5).to(torch.bfloat16)
        act = torch.relu(torch.randn(B, kin, device=dev)).to(torch.bfloat16)
        Y = torch.zeros(B, nout, dtype=torch.bfloat16, device=dev)
        dX = torch.zeros(B, kin, dtype=torch.bfloat16, device=dev)
        dW = torch.zeros(kin, nout, dtype=torch.float32, device=dev)
        for kind in ("fwd", "dgrad", "wgrad"):
            e = _abi.GemmEpilogue()
            e.n_valid, e.n_block, e.pad_one, e.split_k = 1 << 30, 1 << 30, 0.0, 1
            if kind == "fwd" and nout == 64 and os.environ.get("DMVAE_HEAD_SPLIT"):
                # latent head as the engine runs it: fp32 output, k-splits reduced into a cleared buffer
                Yf = torch.zeros(B, nout, dtype=torch.float32, device=dev)
                e.out_dtype, e.act, e.split_k, e.accumulate = F32, _abi.ACT_NONE, int(os.environ["DMVAE_HEAD_SPLIT"]), 1
                M, N, K = B, nout, kin
                call = lambda: lib.dmvae_gemm(ctx, BF16, 0, 0, X.data_ptr(), kin, W.data_ptr(), nout, Yf.data_ptr(), nout,
                                              M, N, K, C.byref(e), st())
                ref = lambda: X.float() @ W.float()
                out = Yf
            elif kind == "fwd":
                e.out_dtype, e.act = BF16, _abi.ACT_RELU
                M, N, K = B, nout, kin
                call = lambda: lib.dmvae_gemm(ctx, BF16, 0, 0, X.data_ptr(), kin, W.data_ptr(), nout, Y.data_ptr(), nout,
                                              M, N, K, C.byref(e), st())
                ref = lambda: torch.relu(X.float() @ W.float())
                out = Y
            elif kind == "dgrad":
                e.out_dtype, e.act = BF16, _abi.ACT_NONE
                e.relu_mask, e.ld_mask = act.data_ptr(), kin
                M, N, K = B, kin, nout
                call = lambda: lib.dmvae_gemm(ctx, BF16, 0, 1, dY.data_ptr(), nout, W.data_ptr(), nout, dX.data_ptr(), kin,
                                              M, N, K, C.byref(e), st())
                ref = lambda: (dY.float() @ W.float().t()) * (act.float() > 0)
                out = dX
            else:
                M, N, K = kin, nout, B
                nkb = (K + 63) // 64
                if M >= 256 and N >= 128 and os.environ.get("DMVAE_GEMM_PAIR") != "0":
                    tiles = ((M + 255) // 256) * ((N + 255) // 256)
                    sk = max(1, min(nkb // 2, 74 // tiles))
                else:
                    tiles = ((M + 127) // 128) * ((N + 127) // 128)
                    sk = max(1, min(nkb, (2 * 148 + tiles - 1) // tiles))
                e.out_dtype, e.act, e.split_k, e.accumulate = F32, _abi.ACT_NONE, sk, 1 if sk > 1 else 0
                call = lambda: lib.dmvae_gemm(ctx, BF16, 1, 0, X.data_ptr(), kin, dY.data_ptr(), nout, dW.data_ptr(), nout,
                                              M, N, K, C.byref(e), st())
                ref = lambda: X.float().t() @ dY.float()
                out = dW
            err = float("nan")
            if args.check:
                out.zero_()
                _abi.check(call())
                torch.cuda.synchronize()
                r = ref()
                err = float((out.float() - r).abs().max() / r.abs().max().clamp_min(1e-6))
                worst = max(worst, err)
            for _ in range(3):
                _abi.check(call())
            torch.cuda.synchronize()
            # the launches are captured into one CUDA graph so that the host's per-call cost stays out of the device time
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(args.iters):
                    _abi.check(call())
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / args.iters
            del g
            flop = 2.0 * M * N * K
            total_us += us
            total_flop += flop
            _print("%-12s %-6s %6d %6d %6d %5d %9.2f %9.1f %10.2e" % (name, kind, M, N, K, e.split_k, us, flop / us * 1e-6, err))
    _print("total %.1f us per step of GEMMs, %.1f TFLOP/s (padded shapes); worst rel err %.2e" %
           (total_us, total_flop / total_us * 1e-6, worst))
    return total_us, total_flop, worst


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    lib = _abi.load()
    ctx = C.c_void_p()
    _abi.check(lib.dmvae_ctx_create(0, C.byref(ctx)))
    _, _, worst = time_gemms(lib, ctx, args.rows, args.iters, args.check, args.only)
    if args.check and not (worst < 2e-2):
        print("CHECK FAILED")
        sys.exit(1)
    lib.dmvae_ctx_destroy(ctx)


if __name__ == "__main__":
    main()
