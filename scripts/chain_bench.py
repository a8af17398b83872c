#!/usr/bin/env python
"""Forward pass (encoder -> heads -> reparameterisation -> decoder) as ONE chained launch vs one launch per layer:
device time per pass (CUDA-graph replay of `iters` passes) and a bitwise comparison of the outputs."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dmvae_b200.engine import Engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    B = args.rows
    eng = Engine(model="dmvae", input_type="binary", input_dim=784, latent_dim=10, n_classes=10, trunk=(500, 500), head=2000,
                 decoder=(2000, 500, 500), name="dmvae", gemm_dtype="bf16", max_rows=B, seed=0)
    X = (torch.rand(B, 784, device="cuda") < 0.13).to(torch.uint8)
    eng.stage_input(X, B)

    def layered():
        eng.encode(B)
        eng.reparam(B, False, False)
        eng.decode(B)
        eng._join()

    def chained():
        eng.forward_chain(B, False)

    outs = {}
    for name, fn in (("layered", layered), ("chained", chained)):
        for t in (eng.decoded, eng.zh, eng.ch, eng.zb, eng.eps):
            t.zero_()
        fn()
        torch.cuda.synchronize()
        outs[name] = [t.clone() for t in (eng.decoded, eng.zh, eng.ch, eng.zb, eng.eps, eng.act["ench"], eng.act["dec3"])]
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(args.iters):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        print("%-8s %8.2f us per forward pass" % (name, e0.elapsed_time(e1) * 1e3 / args.iters))
    names = ["decoded", "zh", "ch", "zb", "eps", "ench", "dec3"]
    ok = True
    for n, a, b in zip(names, outs["layered"], outs["chained"]):
        d = (a.float() - b.float()).abs().max().item()
        print("  max |layered - chained| %-8s %.3e" % (n, d))
        ok = ok and d == 0.0
    print("IDENTICAL" if ok else "DIFFERENT")


if __name__ == "__main__":
    main()
