"""GPU parity of the mixture-of-experts path (BASELINE config 4: `dmoe` / `dvmoe`, models.py:53-221) against the oracle:
the fused expert-mixture kernel (dmvae_moe_fwd_bwd) alone, and whole MoE steps through the engine."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import closed_form as cf
from oracle import reference_graph as rg

pytestmark = pytest.mark.gpu


def relerr(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(np.abs(b).max(), 1e-30))


@pytest.fixture(scope="module")
def lib():
    from dmvae_b200 import _abi
    return _abi.load()


@pytest.fixture(scope="module")
def ctx(lib):
    c = C.c_void_p()
    assert lib.dmvae_ctx_create(0, C.byref(c)) == 0
    yield c
    lib.dmvae_ctx_destroy(c)


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize("classification", [True, False])
@pytest.mark.parametrize("B,E,O,I", [(256, 16, 10, 784), (37, 5, 1, 12), (100, 16, 2, 10)])
def test_moe_kernel_matches_closed_form(lib, ctx, classification, B, E, O, I):
    """dmvae_moe_fwd_bwd on given expert predictions / gate: loss summands, soft output, predicted class, d_pred, d_gate."""
    from dmvae_b200 import _abi
    if classification and O == 1:
        pytest.skip("a single-class softmax is constant: every gradient is exactly zero")
    rs = np.random.RandomState(B + E)
    inp = rs.randn(B, I)
    W = rs.randn(E, O, I) * (0.3 / np.sqrt(I))
    b = rs.randn(O, E) * 0.1
    gate = cf.softmax(rs.randn(B, E))
    if classification:
        Y = np.eye(O)[rs.randint(0, O, size=B)] if O > 1 else np.ones((B, 1))
        c = cf.moe_classification(inp, gate, W, b, Y)
    else:
        Y = rs.randn(B, O)
        c = cf.moe_regression(inp, gate, W, b, Y)
    pred = np.einsum("eoi,bi->boe", W, inp) + b[None]                      # [B,O,E]
    pred_eo = np.ascontiguousarray(pred.transpose(0, 2, 1).reshape(B, E * O)).astype(np.float32)   # column e*O + o
    ldp = (E * O + 63) // 64 * 64
    pd = torch.zeros(B, ldp, device="cuda")
    pd[:, :E * O] = torch.tensor(pred_eo)
    gd = torch.tensor(gate.astype(np.float32), device="cuda")
    Yd = torch.tensor(Y.astype(np.float32), device="cuda")
    ps = torch.zeros(B, 2, device="cuda")
    ysoft = torch.zeros(B, O, device="cuda")
    cls = torch.zeros(B, dtype=torch.int32, device="cuda")
    dpred = torch.full((B, ldp), 7.0, device="cuda")
    dgate = torch.zeros(B, E, device="cuda")
    ma = _abi.MoeArgs()
    ma.classification, ma.rows, ma.E, ma.O = int(classification), B, E, O
    ma.pred, ma.ld_pred = pd.data_ptr(), ldp
    ma.gate, ma.ld_gate = gd.data_ptr(), E
    ma.Y, ma.ldy = Yd.data_ptr(), O
    ma.inv_global_batch = 1.0 / B
    ma.per_sample, ma.y_soft, ma.pred_class = ps.data_ptr(), ysoft.data_ptr(), cls.data_ptr()
    ma.d_pred, ma.dpred_dtype, ma.ld_dpred, ma.dpred_cols = dpred.data_ptr(), 0, ldp, ldp
    ma.d_gate, ma.ld_dgate = dgate.data_ptr(), E
    _abi.check(lib.dmvae_moe_fwd_bwd(ctx, C.byref(ma), stream()))
    torch.cuda.synchronize()
    tol = 1e-4
    assert relerr(ps[:, 0].cpu().numpy(), c["loss_ps"]) < tol
    assert relerr(ysoft.cpu().numpy(), c["Ysoft"] if classification else c["Yhat"]) < tol
    if classification:
        assert np.array_equal(cls.cpu().numpy(), c["pred_class"])
        onehot = np.eye(O)[c["pred_class"]]
        assert np.allclose(ps[:, 1].cpu().numpy(), np.abs(Y - onehot).sum(-1) / 2, atol=1e-6)
    assert relerr(dgate.cpu().numpy(), c["d_gate"]) < tol
    # d_pred [B, E*O] reproduces d_W = d_pred^T . inp and d_b
    dp = dpred[:, :E * O].cpu().numpy().astype(np.float64).reshape(B, E, O)
    assert np.all(dpred[:, E * O:].cpu().numpy() == 0)
    assert relerr(np.einsum("beo,bi->eoi", dp, inp), c["d_W"]) < tol
    assert relerr(dp.sum(0).T, c["d_b"]) < tol


def _moe_oracle(cfg, V, X, Y, eps, scope, classification, lossVAE, featLearn, gemm_round=None):
    """The reference's MoE graph (models.py:53-111, :149-163) on top of the VAE graph, differentiated by autograd."""
    Vt = rg.to_torch(V)
    Xt, et, Yt = (torch.tensor(np.asarray(a), dtype=torch.float64) for a in (X, eps, Y))
    out = rg.forward(cfg, Vt, Xt, et, 1.0, gemm_round=gemm_round)
    inp = torch.relu(out["mean"]) if featLearn else Xt                      # models.py:58-66
    rd = (lambda t: t) if gemm_round is None else gemm_round
    mo = rg.moe_forward(rd(inp), out["cluster_probs"], rd(Vt[scope + "/regression_weights"]), Vt[scope + "/regression_biases"],
                        Yt, classification)
    loss = mo["recon_loss"] + (out["loss"] if lossVAE else 0.0)             # models.py:161-163
    loss.backward()
    grads = {k: (v.grad.numpy().copy() if v.grad is not None else None) for k, v in Vt.items()}
    return float(mo["recon_loss"]), float(mo["error"]), float(out["loss"]), grads


@pytest.mark.parametrize("kind,tier", [("dmoe", "fp32"), ("dvmoe", "fp32"), ("vademoe", "fp32"), ("dmoe", "bf16"),
                                       ("dvmoe", "bf16"), ("vademoe", "bf16")])
def test_moe_step_matches_oracle(kind, tier):
    """DeepMoE (lossVAE=0, featLearn=0, latent 1; runLR_MOE.sh), DeepVariationalMoE (lossVAE=1, featLearn=1; runOur.sh)
    and VaDEMoE (gate = gamma(Z), models.py:265-275): supervised loss, error count, and every parameter gradient of one
    step, 16 classification experts."""
    from dmvae_b200.engine import Engine
    B, D, E, O = 128, 784, 16, 10
    lossVAE, feat = (0, 0) if kind == "dmoe" else (1, 1)
    L = 1 if kind == "dmoe" else 10
    scope = "/".join([kind] * 3)
    moe = dict(n_experts=E, output_dim=O, featLearn=bool(feat), lossVAE=bool(lossVAE), classification=True, scope=scope)
    if kind == "vademoe":
        cfg = rg.GraphConfig.vade(name=kind, input_dim=D, latent_dim=L, n_classes=E)
        eng = Engine(model="vade", input_type="binary", input_dim=D, latent_dim=L, n_classes=E, trunk=(2000, 500, 500), head=0,
                     decoder=(500, 500, 2000), name=kind, gemm_dtype=tier, max_rows=B, moe=moe)
    else:
        cfg = rg.GraphConfig(name=kind, input_dim=D, latent_dim=L, n_classes=E)
        eng = Engine(model="dmvae", input_type="binary", input_dim=D, latent_dim=L, n_classes=E, trunk=(500, 500), head=2000,
                     decoder=(2000, 500, 500), name=kind, gemm_dtype=tier, max_rows=B, moe=moe)
    V = rg.init_variables(cfg, 0)
    rs = np.random.RandomState(11)
    I = L if feat else D
    V[scope + "/regression_weights"] = (rs.randn(E, O, I) * 0.05).astype(np.float32)
    V[scope + "/regression_biases"] = (rs.randn(O, E) * 0.05).astype(np.float32)
    for k in V:
        if k.endswith("bias") or k.endswith("log_vars"):
            V[k] = (rs.randn(*V[k].shape) * 0.05).astype(np.float32)
    eng.load_variables(V)
    X = (rs.uniform(size=(B, D)) < 0.1307).astype(np.float32)
    Y = np.eye(O)[np.arange(B) % O].astype(np.float32)
    eps = rs.randn(B, L).astype(np.float32)
    rnd = rg.bf16_round if tier == "bf16" else None
    sup, err, vae_loss, g = _moe_oracle(cfg, V, X, Y, eps, scope, True, lossVAE, feat, rnd)
    eng.moe_step(torch.tensor(X, device="cuda"), torch.tensor(Y, device="cuda"), B, None, eps=torch.tensor(eps, device="cuda"))
    torch.cuda.synchronize()
    tol = 1e-4 if tier == "fp32" else 2e-2
    ml = eng.moe_loss.cpu().numpy()                       # [sum of the supervised loss summands, error]
    assert abs(ml[0] / B - sup) <= tol * abs(sup), (ml, sup)
    if tier == "fp32":
        assert abs(ml[1] - err) < 0.5, (ml, err)              # error = number of misclassified samples (models.py:101-103)
    if lossVAE:
        assert abs(float(eng.loss_out[3]) - vae_loss) <= tol * abs(vae_loss)
    checked = 0
    for name, gref in g.items():
        if gref is None or name not in eng.vars:
            continue
        if np.abs(gref).max() == 0:
            continue
        got = eng.get_variable(name, grad=True)
        if tier == "fp32":
            assert relerr(got, gref) < tol, name                  # worst element, relative to the tensor's max
        else:                                                     # bf16 tier: Frobenius, as in test_gpu_model._compare
            e = float(np.linalg.norm(got.astype(np.float64) - gref) / np.linalg.norm(gref))
            assert e < 6e-2, (name, e)
        checked += 1
    assert checked >= (10 if kind == "dmoe" else 20)
    eng.close()
