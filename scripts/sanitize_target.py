#!/usr/bin/env python
"""Smallest program that launches every kernel family of libdmvae_b200 once (target of scripts/sanitize.sh):
DMVAE and VaDE steps in both tiers (tcgen05 pair / single-CTA / grouped GEMMs, fp32 SIMT GEMMs, row-tile ELBO, split-tf32
MMA ELBO + streaming reconstruction + MMA reduction, reparameterisation fwd/bwd, split-weight logits fold, Adam and its
background shape), the MoE step, evaluation (argmax + contingency) and the host-memory row gather."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dmvae_b200.engine import Engine  # noqa: E402
from dmvae_b200 import _abi  # noqa: E402


def main():
    rs = np.random.RandomState(0)
    B = 272                                                        # one full 256-row pair tile + a ragged 16-row tile
    X = torch.tensor((rs.uniform(size=(B, 784)) < 0.13).astype(np.uint8), device="cuda")
    for model, tier, L, K, kw in (("dmvae", "bf16", 10, 10, {}), ("dmvae", "fp32", 10, 10, {}), ("vade", "bf16", 64, 50, {}),
                                  ("dmvae", "bf16", 40, 24, {})):  # last: DMVAE beyond the row-tile kernel (K*L > 512)
        trunk, head, dec = ((500, 500), 2000, (2000, 500, 500)) if model == "dmvae" else ((2000, 500, 500), 0, (500, 500, 2000))
        eng = Engine(model=model, input_type="binary", input_dim=784, latent_dim=L, n_classes=K, trunk=trunk, head=head,
                     decoder=dec, name=model, gemm_dtype=tier, max_rows=B)
        opt = eng.optimizer("train", 0.002)
        eng.use_graphs = False
        for _ in range(2):
            eng.train_step(X, B, opt)
        eng.use_graphs = True
        for _ in range(3):                                         # capture + replays: streamed (background) Adam, PDL edges
            eng.train_step(X, B, opt)
        torch.cuda.synchronize()
        assert torch.isfinite(eng.loss_out).all()
        print("%s %s L=%d K=%d: loss %.3f" % (model, tier, L, K, float(eng.loss_out[3])), flush=True)
        eng.close()
    # MoE step + evaluation helpers
    moe = dict(n_experts=16, output_dim=10, featLearn=False, lossVAE=False, classification=True, scope="m/m/m")
    eng = Engine(model="dmvae", input_type="binary", input_dim=784, latent_dim=1, n_classes=16, trunk=(500, 500), head=2000,
                 decoder=(2000, 500, 500), name="dmvae", gemm_dtype="bf16", max_rows=B, moe=moe)
    Y = torch.nn.functional.one_hot(torch.arange(B) % 10, 10).float().cuda()
    opt = eng.optimizer("moe", 0.002)
    eng.moe_step(X, Y, B, opt)
    cls = (torch.arange(B) % 10).int().cuda()
    counts = torch.zeros(16, 16, dtype=torch.int32, device="cuda")
    _abi.check(eng.lib.dmvae_argmax_contingency(eng.ctx, eng.ch.data_ptr(), eng.ch.stride(0), B, 16, cls.data_ptr(), 16,
                                                eng.argmax.data_ptr(), counts.data_ptr(), eng._stream()))
    host = X.cpu().pin_memory()
    idx = torch.randperm(B).int().cuda()
    dst = torch.empty_like(X)
    _abi.check(eng.lib.dmvae_gather_rows(eng.ctx, host.data_ptr(), 784, idx.data_ptr(), dst.data_ptr(), 784, B, 784, eng._stream()))
    torch.cuda.synchronize()
    assert torch.equal(dst.cpu(), host[idx.cpu().long()]) and int(counts.sum()) == B
    print("moe + eval + gather ok", flush=True)
    eng.close()


if __name__ == "__main__":
    main()
