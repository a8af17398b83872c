"""GPU tests of the callers either side of the hot path: the epoch loop with the device-side shuffle gather,
clustering accuracy against the oracle, checkpoint interchange by TF variable name, the MoE training loop, and the
train.py driver for one epoch per model."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import reference_graph as rg


def _clustered(n, seed=0, D=784, K=10):
    rs = np.random.RandomState(seed)
    protos = rs.uniform(size=(K, D)) < 0.2
    cls = np.arange(n) % K
    X = (protos[cls] ^ (rs.uniform(size=(n, D)) < 0.03)).astype(np.float32)
    return X, cls


def _model(gd="fp32", name="dmvae", **kw):
    from dmvae_b200 import base_models, nn
    m = base_models.DeepMixtureVAE(name, "binary", 784, 10, 10, activation=nn.relu, initializer=nn.xavier_initializer, **kw).build_graph()
    m.gemm_dtype = gd
    m.define_train_step(0.002, 100)
    return m


def test_gather_rows_from_pinned_host_and_device():
    """dmvae_gather_rows: dst[i] = src[idx[i]] for 16-byte and 4-byte granular rows, source in pinned host or device memory."""
    import ctypes as C
    from dmvae_b200 import _abi
    lib = _abi.load()
    ctx = C.c_void_p()
    _abi.check(lib.dmvae_ctx_create(0, C.byref(ctx)))
    rs = np.random.RandomState(0)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for n, cols, dt in ((1000, 784, torch.uint8), (333, 10, torch.float32), (64, 3072, torch.float32), (5, 4, torch.uint8)):
        a = torch.from_numpy(rs.randint(0, 255, size=(n, cols))).to(dt)
        idx = torch.from_numpy(rs.permutation(n).astype(np.int32))
        for src in (a.pin_memory(), a.cuda()):
            dst = torch.zeros(n, cols, dtype=dt, device="cuda")
            rb = cols * a.element_size()
            _abi.check(lib.dmvae_gather_rows(ctx, src.data_ptr(), rb, idx.cuda().data_ptr(), dst.data_ptr(), rb, n, rb, st))
            torch.cuda.synchronize()
            assert torch.equal(dst.cpu(), a[idx.long()])
    lib.dmvae_ctx_destroy(ctx)


def test_epoch_with_device_shuffle_equals_epoch_over_preshuffled_rows():
    """train_op's fast path gathers each batch by permutation index on the device; it must give exactly the losses and
    parameters of an epoch over the host array re-ordered with the same permutation (fp32 tier: deterministic)."""
    from dmvae_b200.includes.utils import Dataset
    from dmvae_b200.session import Session
    X, cls = _clustered(1000)
    res = []
    for pre in (False, True):
        np.random.seed(7)
        m = _model("fp32")
        sess = Session()
        data = Dataset((X, cls), batch_size=128)
        eng = m._ensure_engine(sess)
        opt = eng.optimizer("train", 0.002)
        data.begin_epoch()
        if pre:
            host = torch.from_numpy(X[data.perm].astype(np.uint8)).pin_memory()
            loss = eng.run_epoch(host, 128, opt)
        else:
            loss = eng.run_epoch(data.host_tensor(), 128, opt, perm=data.perm)
        res.append((loss, eng.get_variable("dmvae/encoder_network/dense/kernel")))
        eng.close()
    assert res[0][0] == res[1][0]
    assert np.array_equal(res[0][1], res[1][1])


@pytest.mark.parametrize("gd", ["fp32", "bf16"])
def test_get_accuracy_matches_the_oracle(gd):
    """base_models.py:425-432 + utils.py:22-34 on the same weights: argmax of the logits, contingency matrix, Hungarian."""
    from dmvae_b200.includes.utils import Dataset, get_clustering_accuracy
    from dmvae_b200.session import Session
    X, cls = _clustered(3000, seed=3)
    cfg = rg.GraphConfig(model="dmvae", input_dim=784, latent_dim=10, n_classes=10)
    V = rg.init_variables(cfg, 5)
    m = _model(gd)
    m.set_variables(V)
    sess = Session()
    data = Dataset((X, cls), batch_size=100)
    acc = m.get_accuracy(sess, data)
    out, _ = rg.loss_and_grads(cfg, V, X, np.zeros((len(X), 10), np.float32), dtype=torch.float32)
    ref = get_clustering_accuracy(out["logits"], cls)
    if gd == "fp32":
        assert acc == ref
    else:
        assert abs(acc - ref) <= 20.0 / len(X)          # a handful of rows inside the bf16 margin of the trunk may move
    # the reference-shaped entry point agrees with the device path
    lg = np.concatenate([sess.run(m.logits, feed_dict={m.X: X[i:i + 1000]}) for i in range(0, len(X), 1000)])
    assert get_clustering_accuracy(lg, cls) == acc


def test_checkpoint_roundtrip_by_tf_variable_name(tmp_path):
    """train.py:233-258: save -> restore into a fresh model reproduces every variable and the loss; keys are the
    reference's TF variable names, so a TensorFlow run elsewhere can exchange weights with this build."""
    from dmvae_b200.session import Session
    X, _ = _clustered(256, seed=1)
    eps = np.random.RandomState(2).randn(256, 10).astype(np.float32)
    sess = Session()
    a = _model("bf16")
    a.seed = 11
    for _ in range(3):
        sess.run([a.loss, a.train_step], feed_dict={a.X: X, a.epsilon: eps})
    ck = str(tmp_path / "model" / "parameters.ckpt")
    a.save(ck)
    names = set(np.load(ck + ".npz").files)
    cfg = rg.GraphConfig(model="dmvae", input_dim=784, latent_dim=10, n_classes=10)
    assert names == set(n for n, _, _ in rg.variable_specs(cfg))
    b = _model("bf16")
    b.seed = 99                                             # different initialisation: everything must come from the file
    b.restore(ck)
    va, vb = a.get_variables(), b.get_variables()
    for k in va:
        assert np.array_equal(va[k], vb[k]), k
    la = sess.run(a.loss, feed_dict={a.X: X, a.epsilon: eps})
    lb = sess.run(b.loss, feed_dict={b.X: X, b.epsilon: eps})
    assert la == lb


@pytest.mark.parametrize("kind,classification", [("dmoe", True), ("dvmoe", True), ("dvmoe", False), ("vademoe", True)])
def test_moe_train_op_and_get_accuracy(kind, classification):
    """models.py:121-147, :194-221 through the reference API with MEDataset: the loss falls, the returned tuples have the
    reference's meaning, and the replayed-graph epoch equals the session.run-per-batch loop on the first step."""
    from dmvae_b200 import models, nn
    from dmvae_b200.includes.utils import MEDataset
    from dmvae_b200.session import Session
    X, cls = _clustered(1200, seed=4)
    rs = np.random.RandomState(5)
    O = 10 if classification else 2
    if classification:
        Y = np.eye(10)[cls]
    else:
        Wt = rs.randn(10, 784, O) * 0.05
        Y = np.einsum("bd,bdo->bo", X, Wt[cls])
    np.random.seed(3)
    ctor = dict(dmoe=models.DeepMoE, dvmoe=models.DeepVariationalMoE, vademoe=models.VaDEMoE)[kind]
    if kind == "dmoe":
        m = ctor(kind, "binary", 784, O, 10, classification, activation=nn.relu, initializer=nn.xavier_initializer).build_graph()
    else:
        m = ctor(kind, "binary", 784, 10, O, 10, classification, activation=nn.relu, initializer=nn.xavier_initializer).build_graph()
    m.gemm_dtype = "fp32"
    m.define_train_step(0.002, 100)
    sess = Session()
    data = MEDataset((X, cls, Y), batch_size=200)
    hist = [m.train_op(sess, data, 1.0) for _ in range(4)]
    losses = [h[0] for h in hist]
    assert np.isfinite(losses).all() and losses[-1] < losses[0], losses
    loss, batch_acc, lossCls = hist[-1]
    assert np.isfinite(lossCls) and lossCls <= loss + 1e-3 * abs(loss) + 1e-6
    acc, acc_cluster = m.get_accuracy(sess, data)
    if classification:
        assert 0.0 <= acc <= 1.0 and 0.0 <= batch_acc <= 1.0
    else:
        assert acc <= 0.0 and batch_acc <= 0.0
    assert 0.0 <= acc_cluster <= 1.0
    # reference-shaped loop (host noise, session.run per batch) on a plain iterable gives finite values too
    class Plain:
        epoch_len, len = 2, 400
        def get_batches(self):
            yield X[:200], Y[:200], cls[:200]
            yield X[200:400], Y[200:400], cls[200:400]
    l2 = m.train_op(sess, Plain(), 1.0)
    assert np.isfinite(l2[0])


@pytest.mark.parametrize("argv", [
    ["--model", "dmvae"], ["--model", "vade", "--latent_dim", "8"],
    ["--model", "dmoe", "--classification", "--n_experts", "10"],
    ["--model", "dvmoe", "--n_experts", "4", "--output_dim", "2", "--featLearn"],
    ["--model", "dmvae", "--pretrain", "--pretrain_epochs_vae", "1", "--pretrain_epochs_prior", "1"]])
def test_train_main_runs_one_epoch(argv, tmp_path, monkeypatch):
    """The reference's driver (train.py:101-338) with its own flags: default dataset name, one epoch, checkpoint-on-best,
    log file."""
    from dmvae_b200 import train
    monkeypatch.chdir(tmp_path)
    args = train.parser.parse_args(argv + ["--n_epochs", "1", "--data_n", "1500", "--seed", "0", "--batch_size", "100"])
    assert args.dataset == "mnist"                          # the reference's default (synthetic stand-in offline)
    acc = train.main(args)
    assert 0.0 <= abs(acc) <= 1e6
    model = argv[1]
    assert os.path.exists(tmp_path / (model + "_logs.txt"))
    assert os.path.isdir(tmp_path / "saved-models")


@pytest.mark.parametrize("gd,tol", [("fp32", 1e-5), ("bf16", 2e-2)])
def test_decode_and_reconstruct_entry_points(gd, tol):
    """includes/visualization.py:39-46 (reconstructed_X with zero noise) and :83-87 (model.Z fed directly): the sigmoid
    is the output GEMM's epilogue (DMVAE_ACT_SIGMOID); both entry points against the oracle's reconstructed_X."""
    from dmvae_b200.session import Session
    X, _ = _clustered(300, seed=8)
    cfg = rg.GraphConfig(model="dmvae", input_dim=784, latent_dim=10, n_classes=10)
    V = rg.init_variables(cfg, 2)
    m = _model(gd)
    m.set_variables(V)
    sess = Session()
    eps0 = np.zeros((300, 10), np.float32)
    out, _ = rg.loss_and_grads(cfg, V, X, eps0, dtype=torch.float32)
    rec = m.reconstruct(sess, X)
    assert rec.shape == (300, 784) and np.all((rec > 0) & (rec < 1))
    assert np.abs(rec - out["reconstructed_X"]).max() < tol
    dec = sess.run(m.decoded_X, feed_dict={m.X: X, m.epsilon: eps0})
    assert relerr_(dec, out["decoded_X"]) < max(tol, 1e-4)
    # generation: feed Z, fetch reconstructed_X
    Z = np.asarray(out["Z"], np.float32)
    gen = m.decode(sess, Z)
    assert np.abs(gen - out["reconstructed_X"]).max() < tol
    with pytest.raises(ValueError):
        sess.run(m.logits, feed_dict={m.Z: Z})


def relerr_(got, ref):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    return float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30))


def test_bit_packed_rows_expand_to_the_same_batch():
    """dmvae_gather_rows_bits: {0,1} rows kept one bit per element on the host come out as the uint8 batch the byte gather
    gives, for a permuted index list and a ragged width (D % 32 == 16)."""
    import ctypes as C
    from dmvae_b200 import _abi
    from dmvae_b200.includes.utils import _pack_bits
    lib = _abi.load()
    ctx = C.c_void_p()
    _abi.check(lib.dmvae_ctx_create(0, C.byref(ctx)))
    rs = np.random.RandomState(3)
    for N, D, B in ((500, 784, 129), (64, 3072, 64), (40, 16, 7)):
        X = (rs.uniform(size=(N, D)) < 0.3).astype(np.uint8)
        packed = _pack_bits(X)
        assert packed is not None and packed.shape[1] % 16 == 0
        host = torch.from_numpy(packed).pin_memory()
        idx = rs.permutation(N)[:B].astype(np.int32)
        idx_d = torch.tensor(idx, device="cuda")
        out = torch.full((B, D), 7, dtype=torch.uint8, device="cuda")
        _abi.check(lib.dmvae_gather_rows_bits(ctx, host.data_ptr(), host.stride(0), idx_d.data_ptr(), out.data_ptr(), out.stride(0),
                                              B, D, None))
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), X[idx])
    lib.dmvae_ctx_destroy(ctx)
