"""Generates tests/golden/*.npz from the fp64 oracle (oracle/reference_graph.py).

PARITY UNPINNED: the reference has no golden vectors and TensorFlow 1.x cannot run here, so these fixtures pin the
ORACLE (regression) and the CUDA path against it - not the reference's own outputs.  Re-generate with
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_graph as rg  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def make(model, fname, seed):
    if model == "dmvae":
        cfg = rg.GraphConfig(model="dmvae", input_dim=64, latent_dim=4, n_classes=5, trunk=(48, 40), head=56, decoder=(56, 40, 48))
    else:
        cfg = rg.GraphConfig.vade(input_dim=64, latent_dim=6, n_classes=7, trunk=(56, 40, 40), decoder=(40, 40, 56))
    V = rg.init_variables(cfg, seed)
    rs = np.random.RandomState(seed + 1)
    for k in V:
        if k.endswith("bias") or k.endswith("log_vars"):
            V[k] = (rs.randn(*V[k].shape) * 0.1).astype(np.float32)
    B = 24
    X = (rs.uniform(size=(B, cfg.input_dim)) < 0.3).astype(np.float32)
    eps = rs.randn(B, cfg.latent_dim).astype(np.float32)
    out, g = rg.loss_and_grads(cfg, V, X, eps, kl_ratio=0.75)
    d = {"X": X, "eps": eps, "kl_ratio": np.float64(0.75)}
    for k, v in V.items():
        d["var:" + k] = v
    for k in ("recon_ps", "kl_c_ps", "kl_z_ps", "elbo_ps", "loss", "recon_loss", "latent_loss", "mean", "log_var",
              "cluster_probs", "decoded_X"):
        d["out:" + k] = np.asarray(out[k])
    for k in rg.trainable_names(cfg):
        d["grad:" + k] = g[k]
    np.savez_compressed(os.path.join(HERE, fname), **d)
    print(fname, "loss", float(out["loss"]))


if __name__ == "__main__":
    make("dmvae", "dmvae_small.npz", 3)
    make("vade", "vade_small.npz", 4)
