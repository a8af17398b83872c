#!/usr/bin/env python
"""Per-launch table of the training step's GEMMs: entry point, shapes, device time (each launch re-issued 20x in a CUDA
graph, as bench.py's roofline does) and TFLOP/s.          python scripts/gemm_launch_table.py [--config 3]"""
import argparse
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import bench  # noqa: E402
import configs as CFG  # noqa: E402


def shapes(name, args):
    """[(M, N, K)] of one recorded call."""
    from dmvae_b200 import _abi
    if name == "dmvae_gemm":
        return [(args[10], args[11], args[12])]
    if name == "dmvae_linear_fwd":                      # ... rows, n_out_pad, n_in_pad
        return [(args[9], args[10], args[11])]
    if name == "dmvae_linear_dgrad":                    # ... rows, n_in_pad, n_out_pad
        return [(args[11], args[12], args[13])]
    if name == "dmvae_linear_wgrad":                    # dW[n_in, n_out] = X^T dY: contraction over the rows
        return [(args[9], args[10], args[8])]
    if name == "dmvae_gemm_chain":
        arr, n = args[1], args[2]
        return [(arr[i].M, arr[i].N, arr[i].K) for i in range(n)]
    return []


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=3)
    a = ap.parse_args()
    cfg = CFG.CONFIGS[a.config]
    B = cfg["batch"]
    dev = torch.device("cuda", 0)
    eng = CFG.make_engine(cfg, B)
    xs = torch.from_numpy(CFG.synth_inputs(cfg, B)).cuda()
    eng.forward_backward(xs, B, fuse=eng.fuse_recon)
    total, rows = 0.0, []
    rec = bench._Recorder(eng.lib)
    eng.lib = rec
    rec.on = True
    eng.forward_backward(xs, B, fuse=eng.fuse_recon)
    rec.on = False
    eng.lib = rec._lib
    torch.cuda.synchronize()
    from dmvae_b200 import _abi
    st = lambda: C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    print("%-20s %-46s %9s %9s" % ("entry", "M x N x K (per GEMM of the launch)", "us", "TFLOP/s"))
    tot_us = tot_fl = 0.0
    for name, fn, args in rec.calls:
        al = list(args)
        us = bench.graph_time_us(torch, dev, lambda: _abi.check(fn(*(al[:-1] + [st()]))))
        sh = shapes(name, args)
        fl = sum(2.0 * m * n * k for m, n, k in sh)
        tot_us += us
        tot_fl += fl
        print("%-20s %-46s %9.1f %9.1f" % (name.replace("dmvae_", ""), " + ".join("%dx%dx%d" % s for s in sh)[:46], us, fl / us / 1e6))
    print("total: %.1f us, %.1f GFLOP issued (padded shapes), %.1f TFLOP/s; algorithmic %.1f GFLOP" %
          (tot_us, tot_fl / 1e9, tot_fl / tot_us / 1e6, CFG.gemm_flop_per_sample(cfg) * B / 1e9))


if __name__ == "__main__":
    main()
