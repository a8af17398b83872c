"""Hand-derived closed-form ELBO terms and gradients at the fused-kernel boundary
(TEST INFRASTRUCTURE; NumPy float64).

These are the formulas of SURVEY.md section 8(a'), derived from priors.py:104-147, :183-201,
:91-102 and base_models.py:72-93.  They are an *independent* restatement: tests check them
against fp64 autograd of oracle/reference_graph.py, and the CUDA kernel against both.

Boundary: inputs are the tensors the fused ELBO kernel sees (X, decoded logits, mean, log_var,
logits or Z, weights/zeta, prior tables); outputs are per-sample terms, q(c|x), argmax and the
gradients wrt decoded, mean, log_var, logits (or Z), prior means and prior log_vars for
loss = s * sum_b [R_b + r (C_b + Zk_b)].
"""
from __future__ import annotations

import numpy as np

EPS0 = 1e-20


def softmax(x):
    x = x - x.max(axis=-1, keepdims=True)
    e = np.exp(x)
    return e / e.sum(axis=-1, keepdims=True)


def recon_terms(X, decoded, input_type):
    if input_type == "binary":
        R = (np.maximum(decoded, 0) - decoded * X + np.log1p(np.exp(-np.abs(decoded)))).sum(1)
        dR = 1.0 / (1.0 + np.exp(-decoded)) - X
    elif input_type == "real":
        R = 0.5 * ((X - decoded) ** 2).sum(1)
        dR = decoded - X
    else:
        raise NotImplementedError
    return R, dR


def kl_c_terms(q, K):
    C = (q * (np.log(q + EPS0) + np.log(K))).sum(1)
    gC = np.log(q + EPS0) + q / (q + EPS0) + np.log(K)
    return C, gC


def a_matrix(mean, log_var, m, plv):
    """A_bk = sum_l [plv_kl - lv_bl - 1 + (e^{lv_bl} + (mu_bl - m_kl)^2) e^{-plv_kl}]."""
    iv = np.exp(-plv)
    d = mean[:, None, :] - m[None, :, :]
    return (plv[None] - log_var[:, None, :] - 1 + (np.exp(log_var)[:, None, :] + d * d) * iv[None]).sum(-1)


def elbo_dmvae(X, decoded, mean, log_var, logits, m, plv, r=1.0, s=None, input_type="binary"):
    """DMVAE analytic branch (cluster_sample=False, w = q = softmax(logits))."""
    B, K = logits.shape
    s = 1.0 / B if s is None else s
    R, dR = recon_terms(X, decoded, input_type)
    q = softmax(logits)
    C, gC = kl_c_terms(q, K)
    A = a_matrix(mean, log_var, m, plv)
    Zk = 0.5 * (q * A).sum(1)
    iv = np.exp(-plv)
    G = r * (gC + 0.5 * A)
    d_logits = s * q * (G - (q * G).sum(1, keepdims=True))
    wiv = q @ iv                                         # [B,L]
    d_mean = s * r * (mean * wiv - q @ (m * iv))
    d_log_var = s * r * 0.5 * (np.exp(log_var) * wiv - q.sum(1, keepdims=True))
    diff = mean[:, None, :] - m[None]                    # [B,K,L]
    d_m = -s * r * np.einsum("bk,bkl->kl", q, diff * iv[None])
    d_plv = s * r * 0.5 * np.einsum("bk,bkl->kl", q, 1 - (np.exp(log_var)[:, None, :] + diff ** 2) * iv[None])
    return dict(R=R, C=C, Zk=Zk, q=q, argmax=np.argmax(logits, 1), d_decoded=s * dR, d_logits=d_logits,
                d_mean=d_mean, d_log_var=d_log_var, d_means=d_m, d_log_vars=d_plv,
                loss=s * (R + r * (C + Zk)).sum(), elbo=R + r * (C + Zk))


def elbo_dmvae_sampled(X, decoded, mean, log_var, logits, gumbel, tau, m, plv, r=1.0, s=None,
                       input_type="binary"):
    """cluster_sample=True (priors.py:118-128) with w = zeta = softmax((logits+g)/tau) (priors.py:176-178).
    The KL_c term still uses q = softmax(logits)."""
    B, K = logits.shape
    s = 1.0 / B if s is None else s
    R, dR = recon_terms(X, decoded, input_type)
    q = softmax(logits)
    C, gC = kl_c_terms(q, K)
    zeta = softmax((logits + gumbel) / tau)
    mbar = zeta @ m
    pbar = zeta @ plv
    e = np.exp(-pbar)
    dm = mean - mbar
    Zk = 0.5 * (pbar - log_var - 1 + (np.exp(log_var) + dm * dm) * e).sum(1)
    g_mbar = -dm * e
    g_pbar = 0.5 * (1 - (np.exp(log_var) + dm * dm) * e)
    Gw = g_mbar @ m.T + g_pbar @ plv.T                   # dZk/dzeta_k
    d_logits = s * r * q * (gC - (q * gC).sum(1, keepdims=True))
    d_logits = d_logits + (s * r / tau) * zeta * (Gw - (zeta * Gw).sum(1, keepdims=True))
    d_mean = s * r * dm * e
    d_log_var = s * r * 0.5 * (np.exp(log_var) * e - 1)
    d_m = s * r * zeta.T @ g_mbar
    d_plv = s * r * zeta.T @ g_pbar
    return dict(R=R, C=C, Zk=Zk, q=q, zeta=zeta, argmax=np.argmax(logits, 1), d_decoded=s * dR,
                d_logits=d_logits, d_mean=d_mean, d_log_var=d_log_var, d_means=d_m, d_log_vars=d_plv,
                loss=s * (R + r * (C + Zk)).sum(), elbo=R + r * (C + Zk))


def elbo_vade(X, decoded, mean, log_var, Z, m, plv, r=1.0, s=None, input_type="binary"):
    """VaDE: w = gamma = softmax_k(s_k), s_k = -1/2 sum_l[(z_l-m_kl)^2 e^{-plv_kl} + plv_kl]
    (priors.py:91-102); the C term uses probs = gamma (base_models.py:529-536).  Gradients flow through
    gamma into Z (returned as d_Z_gamma; the reparametrisation backward adds it to the decoder's dZ)."""
    B = Z.shape[0]
    K = m.shape[0]
    s = 1.0 / B if s is None else s
    R, dR = recon_terms(X, decoded, input_type)
    iv = np.exp(-plv)
    dz = Z[:, None, :] - m[None]
    sc = -0.5 * ((dz * dz * iv[None]).sum(-1) + plv.sum(-1)[None])
    g = softmax(sc)
    C, gC = kl_c_terms(g, K)
    A = a_matrix(mean, log_var, m, plv)
    Zk = 0.5 * (g * A).sum(1)
    G = r * (gC + 0.5 * A)
    d_s = s * g * (G - (g * G).sum(1, keepdims=True))    # [B,K]
    d_Z = -np.einsum("bk,bkl->bl", d_s, dz * iv[None])
    wiv = g @ iv
    d_mean = s * r * (mean * wiv - g @ (m * iv))
    d_log_var = s * r * 0.5 * (np.exp(log_var) * wiv - g.sum(1, keepdims=True))
    diff = mean[:, None, :] - m[None]
    d_m = -s * r * np.einsum("bk,bkl->kl", g, diff * iv[None]) + np.einsum("bk,bkl->kl", d_s, dz * iv[None])
    d_plv = (s * r * 0.5 * np.einsum("bk,bkl->kl", g, 1 - (np.exp(log_var)[:, None, :] + diff ** 2) * iv[None])
             + np.einsum("bk,bkl->kl", d_s, 0.5 * dz * dz * iv[None] - 0.5))
    return dict(R=R, C=C, Zk=Zk, q=g, argmax=np.argmax(g, 1), d_decoded=s * dR, d_Z_gamma=d_Z, d_s=d_s,
                d_mean=d_mean, d_log_var=d_log_var, d_means=d_m, d_log_vars=d_plv,
                loss=s * (R + r * (C + Zk)).sum(), elbo=R + r * (C + Zk))


def reparam_backward(d_mean_kl, d_log_var_kl, dZ, eps, log_var):
    """priors.py:86-89 backward: Z = mu + exp(lv/2) eps."""
    return d_mean_kl + dZ, d_log_var_kl + 0.5 * dZ * eps * np.exp(log_var / 2)


def moe_classification(inp, gate, W, b, Y, s=None):
    """models.py:76-103, :153-155 forward + closed-form backward.
    Returns loss, Ysoft and gradients wrt pred-level quantities: d_W [E,O,I], d_b [O,E], d_gate [B,E], d_inp [B,I]."""
    B = inp.shape[0]
    s = 1.0 / B if s is None else s
    pred = np.einsum("eoi,bi->boe", W, inp) + b[None]                     # [B,O,E]
    p = softmax(np.transpose(pred, (0, 2, 1)))                            # [B,E,O]
    u = (p * gate[:, :, None]).sum(1)                                     # [B,O]
    S = u.sum(-1, keepdims=True)
    Ys = u / S
    loss_ps = -1000.0 * (Y * np.log(Ys + EPS0)).sum(-1)
    dYs = -1000.0 * s * Y / (Ys + EPS0)                                   # [B,O]
    du = dYs / S - (dYs * u).sum(-1, keepdims=True) / (S * S)             # through the renormalisation
    d_gate = np.einsum("bo,beo->be", du, p)
    dp = du[:, None, :] * gate[:, :, None]                                # [B,E,O]
    dpred_beo = p * (dp - (dp * p).sum(-1, keepdims=True))
    d_W = np.einsum("beo,bi->eoi", dpred_beo, inp)
    d_b = dpred_beo.sum(0).T                                              # [O,E]
    d_inp = np.einsum("beo,eoi->bi", dpred_beo, W)
    return dict(loss=s * loss_ps.sum(), loss_ps=loss_ps, Ysoft=Ys, pred_class=np.argmax(Ys, -1),
                d_W=d_W, d_b=d_b, d_gate=d_gate, d_inp=d_inp)


def moe_regression(inp, gate, W, b, Y, s=None):
    """models.py:105-111, :157-159."""
    B = inp.shape[0]
    s = 1.0 / B if s is None else s
    pred = np.einsum("eoi,bi->boe", W, inp) + b[None]
    Yh = (pred * gate[:, None, :]).sum(-1)
    loss_ps = 0.5 * ((Yh - Y) ** 2).sum(-1)
    dYh = s * (Yh - Y)
    d_gate = np.einsum("bo,boe->be", dYh, pred)
    dpred = dYh[:, :, None] * gate[:, None, :]                            # [B,O,E]
    d_W = np.einsum("boe,bi->eoi", dpred, inp)
    d_b = dpred.sum(0)
    d_inp = np.einsum("boe,eoi->bi", dpred, W)
    return dict(loss=s * loss_ps.sum(), loss_ps=loss_ps, Yhat=Yh, d_W=d_W, d_b=d_b, d_gate=d_gate, d_inp=d_inp)
