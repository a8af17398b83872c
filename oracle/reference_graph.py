"""Op-for-op CPU restatement of the reference TensorFlow graph (TEST INFRASTRUCTURE).

Every function cites the reference file:line it follows (paths relative to
/root/reference/code).  Tensors are torch CPU tensors so that the backward pass
is produced by autograd exactly like ``optimizer.minimize`` does in the
reference (base_models.py:110); dtype is selectable (float64 = truth,
float32 = reference-precision twin and the timed CPU baseline).

PARITY UNPINNED (see oracle/__init__.py): validated by analytic KATs, fp64
autograd vs. hand-derived closed form, and finite differences only.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch


# ----------------------------------------------------------------------------
# configuration
# ----------------------------------------------------------------------------
@dataclass
class GraphConfig:
    """Shapes of one model.  Defaults are the hard-coded reference sizes
    (base_models.py:221-248, :283-285 for DMVAE; :490-499, :541-545 for VaDE)."""
    model: str = "dmvae"            # "dmvae" | "vade"
    input_type: str = "binary"      # "binary" | "real"  (base_models.py:72-85)
    input_dim: int = 784
    latent_dim: int = 10
    n_classes: int = 10
    name: str = "dmvae"
    # DMVAE: trunk (500, 500), head width 2000.  VaDE: encoder (2000, 500, 500).
    trunk: Tuple[int, ...] = (500, 500)
    head: int = 2000
    decoder: Tuple[int, ...] = (2000, 500, 500)
    cluster_sample: bool = False    # base_models.py:271 (always False in the reference)

    @staticmethod
    def vade(**kw) -> "GraphConfig":
        d = dict(model="vade", name="vade", trunk=(2000, 500, 500), head=0,
                 decoder=(500, 500, 2000))
        d.update(kw)
        return GraphConfig(**d)


def xavier_uniform(rng: np.random.RandomState, shape) -> np.ndarray:
    """tf.contrib.layers.xavier_initializer(): U(-a, a), a = sqrt(6/(fan_in+fan_out)),
    fan_in = shape[-2], fan_out = shape[-1] (train.py:161; layers.py:25-28 applies it
    to the (1, out) FullyConnected bias as well, giving a = sqrt(6/(1+out)))."""
    fan_in, fan_out = shape[-2], shape[-1]
    a = math.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-a, a, size=shape).astype(np.float32)


def variable_specs(cfg: GraphConfig) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(tf variable name, shape, init kind) in graph-construction order.
    Names follow the variable scopes of base_models.py:158-302 (DMVAE) and :443-562 (VaDE)."""
    n, D, L, K = cfg.name, cfg.input_dim, cfg.latent_dim, cfg.n_classes
    specs: List[Tuple[str, Tuple[int, ...], str]] = []
    if cfg.model == "dmvae":
        h1, h2 = cfg.trunk
        e = n + "/encoder_network"
        specs += [(e + "/dense/kernel", (D, h1), "xavier"), (e + "/dense/bias", (h1,), "zeros"),
                  (e + "/dense_1/kernel", (h1, h2), "xavier"), (e + "/dense_1/bias", (h2,), "zeros"),
                  (e + "/z/dense/kernel", (h2, cfg.head), "xavier"), (e + "/z/dense/bias", (cfg.head,), "zeros"),
                  (e + "/z/dense_1/kernel", (cfg.head, L), "xavier"), (e + "/z/dense_1/bias", (L,), "zeros"),
                  (e + "/z/dense_2/kernel", (cfg.head, L), "xavier"), (e + "/z/dense_2/bias", (L,), "zeros"),
                  (e + "/c/dense/kernel", (h2, cfg.head), "xavier"), (e + "/c/dense/bias", (cfg.head,), "zeros"),
                  (e + "/c/dense_1/kernel", (cfg.head, K), "xavier"), (e + "/c/dense_1/bias", (K,), "zeros"),
                  # dead head reconstructed_Y_soft, base_models.py:251-253 (never receives a gradient)
                  (e + "/dense_2/kernel", (h2, 10), "xavier"), (e + "/dense_2/bias", (10,), "zeros")]
    elif cfg.model == "vade":
        e = n + "/encoder_network"
        prev = D
        for i, h in enumerate(cfg.trunk):
            specs += [(e + "/layers/layer_%d/weight" % (i + 1), (prev, h), "xavier"),
                      (e + "/layers/layer_%d/bias" % (i + 1), (1, h), "xavier")]
            prev = h
        specs += [(e + "/z/dense/kernel", (prev, L), "xavier"), (e + "/z/dense/bias", (L,), "zeros"),
                  (e + "/z/dense_1/kernel", (prev, L), "xavier"), (e + "/z/dense_1/bias", (L,), "zeros")]
    else:
        raise NotImplementedError(cfg.model)
    # priors.py:58-65
    specs += [(n + "/representation/means", (K, L), "normal"),
              (n + "/representation/log_vars", (K, L), "zeros")]
    d = n + "/decoder_network"
    prev = L
    for i, h in enumerate(cfg.decoder):
        specs += [(d + "/layers/layer_%d/weight" % (i + 1), (prev, h), "xavier"),
                  (d + "/layers/layer_%d/bias" % (i + 1), (1, h), "xavier")]
        prev = h
    specs += [(d + "/dense/kernel", (prev, D), "xavier"), (d + "/dense/bias", (D,), "zeros")]
    return specs


def init_variables(cfg: GraphConfig, seed: int = 0) -> Dict[str, np.ndarray]:
    rng = np.random.RandomState(seed)
    out: Dict[str, np.ndarray] = {}
    for name, shape, kind in variable_specs(cfg):
        if kind == "xavier":
            out[name] = xavier_uniform(rng, shape)
        elif kind == "zeros":
            out[name] = np.zeros(shape, np.float32)
        elif kind == "normal":      # tf.initializers.random_normal: mean 0, std 1 (priors.py:60)
            out[name] = rng.standard_normal(shape).astype(np.float32)
        else:
            raise ValueError(kind)
    return out


def trainable_names(cfg: GraphConfig) -> List[str]:
    """Variables that receive a gradient from ``loss`` (everything except the dead head)."""
    return [n for n, _, _ in variable_specs(cfg) if "/encoder_network/dense_2/" not in n]


# ----------------------------------------------------------------------------
# noise (host side in the reference)
# ----------------------------------------------------------------------------
def sample_gumbel(rng: np.random.RandomState, shape, eps=1e-20):
    """includes/utils.py:17-19."""
    U = rng.uniform(0, 1, shape)
    return -np.log(eps - np.log(U + eps))


def gumbel_from_uniform(U, eps=1e-20):
    return -np.log(eps - np.log(U + eps))


# ----------------------------------------------------------------------------
# graph pieces
# ----------------------------------------------------------------------------
def _dense(x, W, b, act):
    """tf.layers.dense (base_models.py:221-248) and FullyConnected._call (layers.py:30-36):
    matmul + bias, then the activation."""
    y = torch.matmul(x, W) + b
    return torch.relu(y) if act else y


def sigmoid_xent_with_logits(labels, logits):
    """tf.nn.sigmoid_cross_entropy_with_logits stable form:
    max(x,0) - x*z + log(1+exp(-|x|))  (base_models.py:74-79)."""
    return torch.clamp(logits, min=0) - logits * labels + torch.log1p(torch.exp(-torch.abs(logits)))


def inverse_reparametrize(mean, log_var, epsilon):
    """priors.py:86-89."""
    return mean + torch.exp(log_var / 2) * epsilon


def gumbel_softmax(logits, gumbel, temperature):
    """DiscreteFactorial.inverse_reparametrize, priors.py:170-181 (dim = 1)."""
    return torch.softmax((logits + gumbel) / temperature, dim=-1)


def get_cluster_probs(Z, means, log_vars):
    """NormalMixtureFactorial.get_cluster_probs, priors.py:91-102 ([B,K,L] broadcast kept)."""
    Zb = Z[:, None, :]
    m = means[None, :, :]
    lv = log_vars[None, :, :]
    probs = -(torch.sum(torch.square(Zb - m) / torch.exp(lv), dim=-1) + torch.sum(lv, dim=-1)) / 2
    return torch.softmax(probs, dim=-1)


def kl_discrete_per_sample(q_z, n_classes, eps=1e-20):
    """DiscreteFactorial.kl_from_prior summand before the batch mean, priors.py:195-199."""
    res = q_z * (torch.log(q_z + eps) - math.log(1.0 / n_classes))
    return torch.sum(res, dim=1)


def kl_mixture_per_sample(mean, log_var, weights, means, log_vars, cluster_sample):
    """NormalMixtureFactorial.kl_from_prior summand before the batch mean, priors.py:104-147."""
    if cluster_sample:                                  # priors.py:118-128
        prior_mean = torch.matmul(weights, means)
        prior_log_var = torch.matmul(weights, log_vars)
        res = (prior_log_var - log_var - 1 +
               (torch.exp(log_var) + torch.square(mean - prior_mean)) / torch.exp(prior_log_var))
        return 0.5 * torch.sum(res, dim=1)
    prior_means = means[None, :, :]                     # priors.py:130-145
    prior_log_vars = log_vars[None, :, :]
    mean_ = mean[:, None, :]
    log_var_ = log_var[:, None, :]
    res = (prior_log_vars - log_var_ - 1 +
           (torch.exp(log_var_) + torch.square(mean_ - prior_means)) / torch.exp(prior_log_vars))
    res = torch.sum(res, dim=-1)
    res = torch.sum(res * weights, dim=-1)
    return 0.5 * res


def recon_per_sample(X, decoded, input_type):
    """VAE.define_recon_loss summand before the batch mean, base_models.py:72-85."""
    if input_type == "binary":
        return torch.sum(sigmoid_xent_with_logits(X, decoded), dim=1)
    if input_type == "real":
        return 0.5 * torch.sum(torch.square(X - decoded), dim=1)
    raise NotImplementedError


# ----------------------------------------------------------------------------
# the whole graph
# ----------------------------------------------------------------------------
def to_torch(variables: Dict[str, np.ndarray], dtype=torch.float64, requires_grad=True):
    return {k: torch.tensor(np.asarray(v), dtype=dtype).requires_grad_(requires_grad)
            for k, v in variables.items()}


def forward(cfg: GraphConfig, V: Dict[str, torch.Tensor], X, epsilon, kl_ratio=1.0,
            gumbel=None, temperature=None, inv_global_batch=None, gemm_round=None):
    """Forward graph of DeepMixtureVAE.build_graph (base_models.py:218-300, cnn=False branch)
    or VaDE.build_graph (:490-552), plus define_train_loss (:87-93).

    ``gemm_round``: optional callable applied to both GEMM operands (bf16 emulation tier).
    Returns a dict of every named tensor; per-sample terms are the summands of the reference's
    reduce_mean.  ``inv_global_batch`` (default 1/B) replaces the batch mean so data-parallel
    shards can be checked: loss = inv_global_batch * sum_b(...)."""
    n = cfg.name
    rd = (lambda t: t) if gemm_round is None else gemm_round

    def dense(x, wname, bname, act):
        return _dense(rd(x), rd(V[wname]), V[bname], act)

    out = {}
    B = X.shape[0]
    s = (1.0 / B) if inv_global_batch is None else inv_global_batch
    e = n + "/encoder_network"
    means = V[n + "/representation/means"]
    log_vars = V[n + "/representation/log_vars"]
    if cfg.model == "dmvae":
        hidden = dense(X, e + "/dense/kernel", e + "/dense/bias", True)
        hidden = dense(hidden, e + "/dense_1/kernel", e + "/dense_1/bias", True)
        hidden_z = dense(hidden, e + "/z/dense/kernel", e + "/z/dense/bias", True)
        mean = dense(hidden_z, e + "/z/dense_1/kernel", e + "/z/dense_1/bias", False)
        log_var = dense(hidden_z, e + "/z/dense_2/kernel", e + "/z/dense_2/bias", False)
        hidden_c = dense(hidden, e + "/c/dense/kernel", e + "/c/dense/bias", True)
        logits = dense(hidden_c, e + "/c/dense_1/kernel", e + "/c/dense_1/bias", False)
        cluster_probs = torch.softmax(logits, dim=-1)
        Z = inverse_reparametrize(mean, log_var, epsilon)
        if cfg.cluster_sample:
            weights = gumbel_softmax(logits, gumbel, temperature)
        else:
            weights = cluster_probs
        kl_c = kl_discrete_per_sample(cluster_probs, cfg.n_classes)
        out["logits"] = logits
        out["hidden"] = hidden
    else:
        hidden = X
        for i in range(len(cfg.trunk)):
            hidden = dense(hidden, e + "/layers/layer_%d/weight" % (i + 1),
                           e + "/layers/layer_%d/bias" % (i + 1), True)
        mean = dense(hidden, e + "/z/dense/kernel", e + "/z/dense/bias", False)
        log_var = dense(hidden, e + "/z/dense_1/kernel", e + "/z/dense_1/bias", False)
        Z = inverse_reparametrize(mean, log_var, epsilon)
        cluster_probs = get_cluster_probs(Z, means, log_vars)          # base_models.py:526-527
        weights = cluster_probs
        kl_c = kl_discrete_per_sample(cluster_probs, cfg.n_classes)    # "probs" branch, :529-536
    kl_z = kl_mixture_per_sample(mean, log_var, weights, means, log_vars, cfg.cluster_sample)

    d = n + "/decoder_network"
    h = Z
    for i in range(len(cfg.decoder)):
        h = dense(h, d + "/layers/layer_%d/weight" % (i + 1), d + "/layers/layer_%d/bias" % (i + 1), True)
    decoded = dense(h, d + "/dense/kernel", d + "/dense/bias", False)
    recon = recon_per_sample(X, decoded, cfg.input_type)

    out.update(mean=mean, log_var=log_var, cluster_probs=cluster_probs, weights=weights, Z=Z,
               decoded_X=decoded, recon_ps=recon, kl_c_ps=kl_c, kl_z_ps=kl_z,
               reconstructed_X=torch.sigmoid(decoded) if cfg.input_type == "binary" else decoded)
    out["recon_loss"] = s * torch.sum(recon)
    out["latent_loss"] = s * torch.sum(kl_c) + s * torch.sum(kl_z)
    out["loss"] = out["recon_loss"] + kl_ratio * out["latent_loss"]      # base_models.py:91-93
    out["elbo_ps"] = recon + kl_ratio * (kl_c + kl_z)
    return out


def loss_and_grads(cfg, variables: Dict[str, np.ndarray], X, epsilon, kl_ratio=1.0, gumbel=None,
                   temperature=None, dtype=torch.float64, inv_global_batch=None,
                   extra=("decoded_X", "mean", "log_var", "logits", "Z"), loss_key="loss",
                   gemm_round=None):
    """Forward + autograd backward.  Returns (outputs as numpy, grads dict as numpy).
    Gradients wrt intermediates listed in ``extra`` are returned under "d_<name>"."""
    V = to_torch(variables, dtype)
    Xt = torch.tensor(np.asarray(X), dtype=dtype)
    et = torch.tensor(np.asarray(epsilon), dtype=dtype)
    gt = None if gumbel is None else torch.tensor(np.asarray(gumbel), dtype=dtype).reshape(Xt.shape[0], -1)
    out = forward(cfg, V, Xt, et, kl_ratio, gt, temperature, inv_global_batch, gemm_round)
    keep = [k for k in extra if k in out]
    for k in keep:
        out[k].retain_grad()
    out[loss_key].backward()
    grads = {k: (v.grad.detach().numpy().copy() if v.grad is not None else None) for k, v in V.items()}
    for k in keep:
        grads["d_" + k] = None if out[k].grad is None else out[k].grad.detach().numpy().copy()
    outs = {k: v.detach().numpy().copy() for k, v in out.items()}
    return outs, grads


def bf16_round(t: torch.Tensor) -> torch.Tensor:
    """Round-to-nearest-even to bfloat16 and back (GEMM operand rounding of the bf16 tier);
    straight-through for autograd."""
    r = t.detach().to(torch.float32).to(torch.bfloat16).to(t.dtype)
    return t + (r - t.detach())


# ----------------------------------------------------------------------------
# TF-semantics Adam (base_models.py:102-110; tf.train.AdamOptimizer defaults)
# ----------------------------------------------------------------------------
def adam_tf_step(theta, g, m, v, t, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """One tf.train.AdamOptimizer update (epsilon OUTSIDE the bias-corrected root):
    lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
    theta -= lr_t * m / (sqrt(v) + eps).  Arrays are updated in place; t starts at 1."""
    lr_t = lr * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
    m *= beta1
    m += (1.0 - beta1) * g
    v *= beta2
    v += (1.0 - beta2) * g * g
    theta -= lr_t * m / (np.sqrt(v) + eps)
    return theta, m, v


# ----------------------------------------------------------------------------
# MoE expert block (models.py:53-111, :149-163)
# ----------------------------------------------------------------------------
def moe_forward(inp, gate, W, b, Y, classification, inv_global_batch=None):
    """inp [B,I]; gate [B,E] (= vae.cluster_probs, models.py:74); W [E,O,I]; b [O,E]; Y [B,O].
    expert_predictions[b,o,e] = sum_i W[e,o,i] inp[b,i] + b[o,e]  (models.py:76-81)."""
    B = inp.shape[0]
    s = (1.0 / B) if inv_global_batch is None else inv_global_batch
    # tf.transpose(matmul(W[E,O,I], tile(inp^T)[E,I,B])) -> [B,O,E]
    pred = torch.einsum("eoi,bi->boe", W, inp) + b[None, :, :]
    out = {"pred": pred}
    if classification:
        p = torch.softmax(pred.permute(0, 2, 1), dim=-1)                 # [B,E,O], models.py:84-86
        unnorm = torch.sum(p * gate[:, :, None], dim=1)                   # [B,O]
        Ysoft = unnorm / torch.sum(unnorm, dim=-1, keepdim=True)          # models.py:91-93
        out["reconstructed_Y_soft"] = Ysoft
        out["pred_class"] = torch.argmax(Ysoft, dim=-1)
        onehot = torch.nn.functional.one_hot(out["pred_class"], Y.shape[1]).to(Y.dtype)
        out["error"] = torch.sum(torch.abs(Y - onehot)) / 2                # models.py:101-103
        out["recon_ps"] = -1000.0 * torch.sum(Y * torch.log(Ysoft + 1e-20), dim=-1)
        out["recon_loss"] = s * torch.sum(out["recon_ps"])                # models.py:153-155
    else:
        Yhat = torch.sum(pred * gate[:, None, :], dim=-1)                 # models.py:105-107
        out["reconstructed_Y"] = Yhat
        O = Y.shape[1]
        out["error"] = torch.mean(torch.square(Yhat - Y)) * O             # models.py:109-111
        out["recon_ps"] = 0.5 * torch.sum(torch.square(Yhat - Y), dim=-1)  # = 0.5*mean(.)*O per sample
        out["recon_loss"] = s * torch.sum(out["recon_ps"])                # models.py:157-159
    return out
