"""Timed CPU restatement of the reference training loop (TEST / BASELINE INFRASTRUCTURE).

TensorFlow 1.x cannot be imported here, so the "reference CPU path" that bench.py times is this
op-for-op fp32 PyTorch-CPU restatement of ``VAE.train_op`` (base_models.py:112-132): Python-side
batching like ``Dataset.get_batches`` (includes/utils.py:449-463), host NumPy noise
(priors.py:67-68, utils.py:17-19), unfused graph with the [B,K,L] broadcasts materialised
(priors.py:131-145), autograd backward, one TF-semantics Adam update per variable
(base_models.py:102-110).  kind = "port".
"""
from __future__ import annotations

import math
import time
from typing import Dict

import numpy as np
import torch

from . import reference_graph as rg


class Dataset:
    """includes/utils.py:428-466 (row-by-row Python batching kept on purpose)."""

    def __init__(self, data, batch_size=100, shuffle=True, rng=None):
        data, classes = data
        self.rng = rng or np.random.RandomState(1234)
        self.data = np.copy(data)
        self.classes = np.copy(classes)
        self.batch_size = batch_size
        self.shuffle = shuffle
        self.data_dim = self.data.shape[1]
        self.epoch_len = int(math.ceil(len(self.data) / batch_size))
        if shuffle:
            idx = self.rng.permutation(len(self.data))
            self.data = self.data[idx]
            self.classes = self.classes[idx]

    def get_batches(self):
        if self.shuffle:
            idx = self.rng.permutation(len(self.data))
            self.data = self.data[idx]
            self.classes = self.classes[idx]
        batch = []
        for row in self.data:
            batch.append(row)
            if len(batch) == self.batch_size:
                yield np.array(batch)
                batch = []
        if len(batch) > 0:
            yield np.array(batch)


class CpuTrainer:
    """One reference-style training step on the host."""

    def __init__(self, cfg: rg.GraphConfig, variables: Dict[str, np.ndarray], lr=0.002, seed=2,
                 dtype=torch.float32):
        self.cfg = cfg
        self.dtype = dtype
        self.names = rg.trainable_names(cfg)
        self.V = {k: torch.tensor(v, dtype=dtype) for k, v in variables.items()}
        for k in self.names:
            self.V[k].requires_grad_(True)
        self.m = {k: torch.zeros_like(self.V[k]) for k in self.names}
        self.v = {k: torch.zeros_like(self.V[k]) for k in self.names}
        self.t = 0
        self.lr = lr
        self.rng = np.random.RandomState(seed)

    def step(self, batch: np.ndarray, kl_ratio=1.0) -> float:
        cfg = self.cfg
        B = len(batch)
        # host noise every step, as sample_reparametrization_variables does (base_models.py:44-56)
        eps = self.rng.randn(B, cfg.latent_dim)
        gum = rg.sample_gumbel(self.rng, (B, 1, cfg.n_classes))     # fed but unconsumed by DMVAE
        X = torch.tensor(batch, dtype=self.dtype)
        e = torch.tensor(eps, dtype=self.dtype)
        g = torch.tensor(gum, dtype=self.dtype).reshape(B, -1)
        out = rg.forward(cfg, self.V, X, e, kl_ratio, g, 1.0)
        loss = out["loss"]
        grads = torch.autograd.grad(loss, [self.V[k] for k in self.names], allow_unused=True)
        self.t += 1
        b1, b2, eps_a = 0.9, 0.999, 1e-8
        lr_t = self.lr * math.sqrt(1.0 - b2 ** self.t) / (1.0 - b1 ** self.t)
        with torch.no_grad():
            for k, gk in zip(self.names, grads):                     # one ApplyAdam per variable
                if gk is None:
                    continue
                self.m[k].mul_(b1).add_(gk, alpha=1 - b1)
                self.v[k].mul_(b2).addcmul_(gk, gk, value=1 - b2)
                self.V[k].sub_(lr_t * self.m[k] / (self.v[k].sqrt() + eps_a))
        return float(loss)


class MoeCpuTrainer:
    """`dmoe` training step on the host (models.py:53-111, :149-163, :194-221 with DeepMoE's lossVAE=0, featLearn=0):
    only the sub-graph that `session.run([error, loss, train_step, recon_loss])` evaluates - trunk, c-head, softmax gate,
    tiled-input expert matmul, gated mixture, NLL x 1000 - differentiated by autograd, one TF-semantics Adam per variable."""

    def __init__(self, cfg: rg.GraphConfig, variables: Dict[str, np.ndarray], n_experts, output_dim, lr=0.002, seed=2,
                 dtype=torch.float32):
        self.cfg, self.dtype, self.E, self.O = cfg, dtype, n_experts, output_dim
        rs = np.random.RandomState(seed)
        e = cfg.name + "/encoder_network"
        keep = [e + "/dense/kernel", e + "/dense/bias", e + "/dense_1/kernel", e + "/dense_1/bias", e + "/c/dense/kernel",
                e + "/c/dense/bias", e + "/c/dense_1/kernel", e + "/c/dense_1/bias"]
        self.V = {k: torch.tensor(variables[k], dtype=dtype).requires_grad_(True) for k in keep}
        self.V["W"] = torch.tensor(rs.randn(n_experts, output_dim, cfg.input_dim), dtype=dtype).requires_grad_(True)   # models.py:53-72
        self.V["b"] = torch.zeros(output_dim, n_experts, dtype=dtype).requires_grad_(True)
        self.names = list(self.V.keys())
        self.m = {k: torch.zeros_like(v) for k, v in self.V.items()}
        self.v = {k: torch.zeros_like(v) for k, v in self.V.items()}
        self.t, self.lr = 0, lr
        self.rng = rs

    def step(self, batch, labels) -> float:
        cfg, V = self.cfg, self.V
        e = cfg.name + "/encoder_network"
        B = len(batch)
        self.rng.randn(B, cfg.latent_dim)                                # noise is drawn and fed every step (models.py:200-207)
        rg.sample_gumbel(self.rng, (B, 1, cfg.n_classes))
        X = torch.tensor(batch, dtype=self.dtype)
        Y = torch.tensor(labels, dtype=self.dtype)
        h = torch.relu(X @ V[e + "/dense/kernel"] + V[e + "/dense/bias"])
        h = torch.relu(h @ V[e + "/dense_1/kernel"] + V[e + "/dense_1/bias"])
        hc = torch.relu(h @ V[e + "/c/dense/kernel"] + V[e + "/c/dense/bias"])
        gate = torch.softmax(hc @ V[e + "/c/dense_1/kernel"] + V[e + "/c/dense_1/bias"], dim=-1)
        mo = rg.moe_forward(X, gate, V["W"], V["b"], Y, True)
        loss = mo["recon_loss"]
        grads = torch.autograd.grad(loss, [V[k] for k in self.names])
        self.t += 1
        b1, b2, eps_a = 0.9, 0.999, 1e-8
        lr_t = self.lr * math.sqrt(1.0 - b2 ** self.t) / (1.0 - b1 ** self.t)
        with torch.no_grad():
            for k, gk in zip(self.names, grads):
                self.m[k].mul_(b1).add_(gk, alpha=1 - b1)
                self.v[k].mul_(b2).addcmul_(gk, gk, value=1 - b2)
                V[k].sub_(lr_t * self.m[k] / (self.v[k].sqrt() + eps_a))
        return float(loss)


def synthetic_binarised(n, dim, seed=1, p=0.1307):
    """X[b,d] = 1{u < 0.1307}, labels b mod 10."""
    rng = np.random.RandomState(seed)
    X = (rng.uniform(size=(n, dim)) < p).astype(np.float32)
    y = (np.arange(n) % 10).astype(np.int32)
    return X, y


def time_training(cfg: rg.GraphConfig, batch_size: int, n_steps: int, warmup: int, threads: int,
                  seed_data=1, seed_w=0, seed_noise=2, binarised=True, moe=None):
    """Returns dict(samples_per_s, ms_per_step (median), p10, p90, cores).  moe = dict(n_experts, output_dim): the
    `dmoe` step (MoeCpuTrainer) instead of the VAE step."""
    torch.set_num_threads(threads)
    n = batch_size * (n_steps + warmup)
    if binarised:
        X, y = synthetic_binarised(n, cfg.input_dim, seed_data)
    else:                                            # CIFAR-shaped soft targets (includes/utils.py:204-210)
        rng = np.random.RandomState(seed_data)
        X = (rng.randint(0, 256, size=(n, cfg.input_dim)) / 255.0).astype(np.float32)
        y = (np.arange(n) % 10).astype(np.int32)
    data = Dataset((X, y), batch_size=batch_size, shuffle=True)
    if moe is None:
        tr = CpuTrainer(cfg, rg.init_variables(cfg, seed_w), seed=seed_noise)
        do_step = lambda b, lo: tr.step(b)
    else:
        tr = MoeCpuTrainer(cfg, rg.init_variables(cfg, seed_w), moe["n_experts"], moe["output_dim"], seed=seed_noise)
        onehot = np.eye(moe["output_dim"], dtype=np.float32)
        do_step = lambda b, lo: tr.step(b, onehot[(np.arange(lo, lo + len(b)) % moe["output_dim"])])
    times = []
    it = data.get_batches()
    t_prev = time.perf_counter()
    for i, batch in enumerate(it):
        do_step(batch, i * batch_size)
        t_now = time.perf_counter()
        if i >= warmup:
            times.append(t_now - t_prev)      # includes the Python batching of this batch
        t_prev = t_now
        if len(times) >= n_steps:
            break
    times = np.array(times)
    med = float(np.median(times))
    return dict(samples_per_s=batch_size / med, ms_per_step=med * 1e3,
                p10_ms=float(np.percentile(times, 10)) * 1e3, p90_ms=float(np.percentile(times, 90)) * 1e3,
                mean_samples_per_s=batch_size * len(times) / float(times.sum()),
                cores=threads, steps=len(times))
