"""Known-answer tests that pin the oracle: analytic values derived
directly from the reference formulas, fp64 autograd vs. the hand-derived closed form, and
central finite differences.  CPU only."""
import math

import numpy as np
import pytest
import torch

from oracle import closed_form as cf
from oracle import philox
from oracle import reference_graph as rg

T = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64)


def test_bce_zero_logits():
    X = np.random.RandomState(0).uniform(size=(3, 784))
    r = rg.recon_per_sample(T(X), T(np.zeros((3, 784))), "binary").numpy()
    assert np.allclose(r, 784 * math.log(2.0), rtol=1e-12)
    assert abs(784 * math.log(2.0) - 543.4274) < 1e-3


def test_bce_grad_is_sigmoid_minus_x():
    rs = np.random.RandomState(1)
    X, lg = rs.uniform(size=(4, 7)), rs.randn(4, 7) * 3
    l = T(lg).requires_grad_(True)
    rg.recon_per_sample(T(X), l, "binary").sum().backward()
    assert np.allclose(l.grad.numpy(), 1 / (1 + np.exp(-lg)) - X, atol=1e-12)
    _, dR = cf.recon_terms(X, lg, "binary")
    assert np.allclose(l.grad.numpy(), dR, atol=1e-12)


def test_real_recon():
    rs = np.random.RandomState(1)
    X, d = rs.randn(4, 7), rs.randn(4, 7)
    r = rg.recon_per_sample(T(X), T(d), "real").numpy()
    assert np.allclose(r, 0.5 * ((X - d) ** 2).sum(1))
    with pytest.raises(NotImplementedError):
        rg.recon_per_sample(T(X), T(d), "other")


def test_kl_c_uniform_and_peaked():
    K = 10
    q = T(np.full((2, K), 1.0 / K))
    v = rg.kl_discrete_per_sample(q, K).numpy()
    assert np.allclose(v, math.log(1.0 / K + 1e-20) - math.log(1.0 / K), atol=1e-15)
    lg = np.zeros((1, K)); lg[0, 3] = 100.0
    v = rg.kl_discrete_per_sample(torch.softmax(T(lg), -1), K).numpy()
    assert abs(v[0] - math.log(K)) < 1e-9


def test_kl_z_identity_and_scalar():
    rs = np.random.RandomState(2)
    K, L = 4, 3
    m = np.tile(rs.randn(1, L), (K, 1)); plv = np.tile(rs.randn(1, L), (K, 1))
    w = cf.softmax(rs.randn(5, K))
    v = rg.kl_mixture_per_sample(T(np.tile(m[:1], (5, 1))), T(np.tile(plv[:1], (5, 1))), T(w), T(m), T(plv), False)
    assert np.allclose(v.numpy(), 0, atol=1e-12)
    v = rg.kl_mixture_per_sample(T([[1.0]]), T([[0.0]]), T([[1.0]]), T([[0.0]]), T([[0.0]]), False)
    assert abs(float(v) - 0.5) < 1e-15


def test_kl_z_sampled_onehot_equals_analytic():
    rs = np.random.RandomState(3)
    B, K, L = 6, 5, 4
    mu, lv, m, plv = rs.randn(B, L), rs.randn(B, L) * .3, rs.randn(K, L), rs.randn(K, L) * .3
    w = np.eye(K)[rs.randint(0, K, B)]
    a = rg.kl_mixture_per_sample(T(mu), T(lv), T(w), T(m), T(plv), False).numpy()
    b = rg.kl_mixture_per_sample(T(mu), T(lv), T(w), T(m), T(plv), True).numpy()
    assert np.allclose(a, b, rtol=1e-12)


def test_vade_cluster_probs_kat():
    rs = np.random.RandomState(4)
    K, L = 5, 3
    m = rs.randn(K, L)
    j = 2
    g = rg.get_cluster_probs(T(m[j:j + 1]), T(m), T(np.zeros((K, L)))).numpy()[0]
    ref = cf.softmax(-0.5 * ((m[j][None] - m) ** 2).sum(1))
    assert np.allclose(g, ref, rtol=1e-12)


def test_reparam_eps_zero_and_gumbel_kat():
    mu, lv = T([[1.0, -2.0]]), T([[0.3, 0.1]])
    assert torch.equal(rg.inverse_reparametrize(mu, lv, torch.zeros_like(mu)), mu)
    assert abs(float(rg.gumbel_from_uniform(np.array(0.5))) - 0.3665129) < 1e-6
    lg = T([[0.1, 0.5, -0.2]]); g = T([[0.0, -1.0, 2.0]])
    z = rg.gumbel_softmax(lg, g, 1e-3).numpy()[0]
    assert np.argmax(z) == np.argmax((lg + g).numpy()[0]) and z.max() > 0.999999


def test_adam_tf_first_step():
    rs = np.random.RandomState(5)
    th0 = rs.randn(50); g = rs.randn(50)
    g[np.abs(g) < 1e-3] = 0.5
    th = th0.copy(); m = np.zeros(50); v = np.zeros(50)
    rg.adam_tf_step(th, g, m, v, 1, 0.002)
    assert np.allclose(th, th0 - 0.002 * np.sign(g), atol=1e-7)
    # epsilon placement: lr_t*m/(sqrt(v)+eps) with UNcorrected m, v
    lr_t = 0.002 * math.sqrt(1 - 0.999) / (1 - 0.9)
    assert np.allclose(th, th0 - lr_t * (0.1 * g) / (np.sqrt(0.001 * g * g) + 1e-8), rtol=1e-12)


def _small_cfg(model, **kw):
    if model == "dmvae":
        return rg.GraphConfig(model="dmvae", input_dim=12, latent_dim=3, n_classes=4, trunk=(7, 6), head=9,
                              decoder=(8, 5, 6), **kw)
    return rg.GraphConfig.vade(input_dim=12, latent_dim=3, n_classes=4, trunk=(9, 7, 6), decoder=(5, 6, 8), **kw)


def _rand_vars(cfg, seed):
    V = rg.init_variables(cfg, seed)
    rs = np.random.RandomState(seed + 100)
    for k in V:                      # make biases / prior log_vars non-trivial
        if k.endswith("bias") or k.endswith("log_vars"):
            V[k] = (rs.randn(*V[k].shape) * 0.2).astype(np.float32)
    return V


@pytest.mark.parametrize("input_type", ["binary", "real"])
@pytest.mark.parametrize("r", [1.0, 0.3])
def test_closed_form_dmvae_matches_autograd(input_type, r):
    cfg = _small_cfg("dmvae", input_type=input_type)
    V = _rand_vars(cfg, 0)
    rs = np.random.RandomState(7)
    B = 5
    X = rs.uniform(size=(B, cfg.input_dim)); eps = rs.randn(B, cfg.latent_dim)
    out, g = rg.loss_and_grads(cfg, V, X, eps, kl_ratio=r)
    n = cfg.name
    c = cf.elbo_dmvae(X, out["decoded_X"], out["mean"], out["log_var"], out["logits"],
                      V[n + "/representation/means"].astype(np.float64),
                      V[n + "/representation/log_vars"].astype(np.float64), r=r, input_type=input_type)
    assert np.allclose(c["R"], out["recon_ps"], rtol=1e-12)
    assert np.allclose(c["C"], out["kl_c_ps"], rtol=1e-10, atol=1e-14)
    assert np.allclose(c["Zk"], out["kl_z_ps"], rtol=1e-12)
    assert abs(c["loss"] - out["loss"]) < 1e-12 * abs(out["loss"])
    assert np.allclose(c["d_decoded"], g["d_decoded_X"], atol=1e-14)
    assert np.allclose(c["d_logits"], g["d_logits"], atol=1e-14)
    assert np.allclose(c["d_means"], g[n + "/representation/means"], atol=1e-14)
    assert np.allclose(c["d_log_vars"], g[n + "/representation/log_vars"], atol=1e-14)
    # d_mean / d_log_var: autograd holds the KL part + the decoder part through Z; add the reparam backward
    dm, dlv = cf.reparam_backward(c["d_mean"], c["d_log_var"], g["d_Z"], eps, out["log_var"])
    assert np.allclose(dm, g["d_mean"], atol=1e-14)
    assert np.allclose(dlv, g["d_log_var"], atol=1e-14)


def test_closed_form_sampled_matches_autograd():
    cfg = _small_cfg("dmvae", cluster_sample=True)
    V = _rand_vars(cfg, 1)
    rs = np.random.RandomState(8)
    B = 6
    X = (rs.uniform(size=(B, cfg.input_dim)) < .3) * 1.0; eps = rs.randn(B, cfg.latent_dim)
    gum = rg.sample_gumbel(rs, (B, 1, cfg.n_classes)); tau = 0.7
    out, g = rg.loss_and_grads(cfg, V, X, eps, kl_ratio=0.8, gumbel=gum, temperature=tau)
    n = cfg.name
    c = cf.elbo_dmvae_sampled(X, out["decoded_X"], out["mean"], out["log_var"], out["logits"], gum.reshape(B, -1),
                              tau, V[n + "/representation/means"].astype(np.float64),
                              V[n + "/representation/log_vars"].astype(np.float64), r=0.8)
    assert np.allclose(c["Zk"], out["kl_z_ps"], rtol=1e-12)
    assert np.allclose(c["zeta"], out["weights"], rtol=1e-12)
    assert np.allclose(c["d_logits"], g["d_logits"], atol=1e-14)
    assert np.allclose(c["d_means"], g[n + "/representation/means"], atol=1e-14)
    assert np.allclose(c["d_log_vars"], g[n + "/representation/log_vars"], atol=1e-14)
    dm, dlv = cf.reparam_backward(c["d_mean"], c["d_log_var"], g["d_Z"], eps, out["log_var"])
    assert np.allclose(dm, g["d_mean"], atol=1e-14) and np.allclose(dlv, g["d_log_var"], atol=1e-14)


def test_closed_form_vade_matches_autograd():
    cfg = _small_cfg("vade")
    V = _rand_vars(cfg, 2)
    rs = np.random.RandomState(9)
    B = 5
    X = rs.uniform(size=(B, cfg.input_dim)); eps = rs.randn(B, cfg.latent_dim)
    n = cfg.name
    m = V[n + "/representation/means"].astype(np.float64); plv = V[n + "/representation/log_vars"].astype(np.float64)
    # autograd with Z as a leaf-like retained tensor: d_Z holds decoder + gamma paths
    out, g = rg.loss_and_grads(cfg, V, X, eps, kl_ratio=0.9)
    c = cf.elbo_vade(X, out["decoded_X"], out["mean"], out["log_var"], out["Z"], m, plv, r=0.9)
    assert np.allclose(c["q"], out["cluster_probs"], rtol=1e-12)
    assert np.allclose(c["C"], out["kl_c_ps"], rtol=1e-10, atol=1e-14)
    assert np.allclose(c["Zk"], out["kl_z_ps"], rtol=1e-12)
    assert np.allclose(c["d_means"], g[n + "/representation/means"], atol=1e-14)
    assert np.allclose(c["d_log_vars"], g[n + "/representation/log_vars"], atol=1e-14)
    # decoder-only dZ: recompute by autograd of the recon loss alone
    _, g2 = rg.loss_and_grads(cfg, V, X, eps, kl_ratio=0.9, loss_key="recon_loss")
    dZ_total = g2["d_Z"] + c["d_Z_gamma"]
    assert np.allclose(dZ_total, g["d_Z"], atol=1e-14)
    dm, dlv = cf.reparam_backward(c["d_mean"], c["d_log_var"], dZ_total, eps, out["log_var"])
    assert np.allclose(dm, g["d_mean"], atol=1e-14) and np.allclose(dlv, g["d_log_var"], atol=1e-14)


def test_finite_differences_on_oracle():
    cfg = _small_cfg("dmvae")
    V = _rand_vars(cfg, 3)
    rs = np.random.RandomState(10)
    X = rs.uniform(size=(4, cfg.input_dim)); eps = rs.randn(4, cfg.latent_dim)
    _, g = rg.loss_and_grads(cfg, V, X, eps)
    V64 = {k: v.astype(np.float64) for k, v in V.items()}
    for name in [cfg.name + "/representation/means", cfg.name + "/encoder_network/c/dense_1/kernel",
                 cfg.name + "/decoder_network/layers/layer_1/bias"]:
        idx = tuple(rs.randint(0, s) for s in V[name].shape)
        h = 1e-6
        vals = []
        for sgn in (+1, -1):
            Vp = {k: v.copy() for k, v in V64.items()}
            Vp[name][idx] += sgn * h
            o = rg.forward(cfg, rg.to_torch(Vp, requires_grad=False), T(X), T(eps))
            vals.append(float(o["loss"]))
        fd = (vals[0] - vals[1]) / (2 * h)
        assert abs(fd - g[name][idx]) < 1e-6 * max(1.0, abs(fd))


def test_dp_shard_gradient_sum_equals_full_batch():
    cfg = _small_cfg("dmvae")
    V = _rand_vars(cfg, 4)
    rs = np.random.RandomState(11)
    B = 8
    X = rs.uniform(size=(B, cfg.input_dim)); eps = rs.randn(B, cfg.latent_dim)
    _, g = rg.loss_and_grads(cfg, V, X, eps)
    parts = [rg.loss_and_grads(cfg, V, X[i::2], eps[i::2], inv_global_batch=1.0 / B)[1] for i in range(2)]
    for k in rg.trainable_names(cfg):
        assert np.allclose(parts[0][k] + parts[1][k], g[k], atol=1e-14)


def test_moe_closed_form_matches_autograd():
    rs = np.random.RandomState(12)
    B, I, E, O = 6, 5, 4, 3
    inp, W, b = rs.randn(B, I), rs.randn(E, O, I), rs.randn(O, E) * .1
    gate = cf.softmax(rs.randn(B, E))
    Y = np.eye(O)[rs.randint(0, O, B)]
    ts = [T(a).requires_grad_(True) for a in (inp, gate, W, b)]
    o = rg.moe_forward(ts[0], ts[1], ts[2], ts[3], T(Y), True)
    o["recon_loss"].backward()
    c = cf.moe_classification(inp, gate, W, b, Y)
    assert abs(c["loss"] - float(o["recon_loss"])) < 1e-10
    for got, t in zip((c["d_inp"], c["d_gate"], c["d_W"], c["d_b"]), ts):
        assert np.allclose(got, t.grad.numpy(), atol=1e-11)
    # single expert: Ysoft = softmax(pred), loss = 1000*CE   
    o1 = rg.moe_forward(T(inp), T(np.ones((B, 1))), T(W[:1]), T(b[:, :1]), T(Y), True)
    sm = cf.softmax(np.einsum("oi,bi->bo", W[0], inp) + b[:, 0][None])
    assert np.allclose(o1["reconstructed_Y_soft"].numpy(), sm, rtol=1e-12)
    assert np.allclose(o1["recon_ps"].numpy(), -1000 * np.log((sm * Y).sum(1) + 1e-20), rtol=1e-12)
    # regression
    Yr = rs.randn(B, O)
    ts = [T(a).requires_grad_(True) for a in (inp, gate, W, b)]
    o = rg.moe_forward(ts[0], ts[1], ts[2], ts[3], T(Yr), False)
    o["recon_loss"].backward()
    c = cf.moe_regression(inp, gate, W, b, Yr)
    assert abs(c["loss"] - float(o["recon_loss"])) < 1e-12
    for got, t in zip((c["d_inp"], c["d_gate"], c["d_W"], c["d_b"]), ts):
        assert np.allclose(got, t.grad.numpy(), atol=1e-12)


def test_philox_random123_known_answers():
    """Random123 v1.14 kat_vectors: philox4x32 10 rounds."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, exp in kat:
        got = philox.philox4x32_10(*[np.array([c], np.uint64) for c in ctr], key[0], key[1])
        assert tuple(int(g[0]) for g in got) == exp


def test_philox_noise_statistics_and_offsets():
    e = philox.normal(4096, 10, seed=2, step=0)
    assert abs(e.mean()) < 0.02 and abs(e.std() - 1) < 0.02
    # global row offsets reproduce the 1-GPU stream on shards
    assert np.array_equal(philox.normal(64, 10, 2, 5, row_offset=32), philox.normal(96, 10, 2, 5)[32:])
    g = philox.gumbel(4096, 10, seed=2, step=0)
    assert abs(g.mean() - 0.5772) < 0.03
