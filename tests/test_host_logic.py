"""CPU tests: the C-ABI library loads and exports every symbol include/dmvae_b200.h declares (no compute calls), the
padded parameter layout, the host-side mirror of the reference API (datasets, accuracy, priors' assertions, train.py
flags) and the oracle against the committed golden fixtures."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_library_exports_every_declared_symbol():
    from dmvae_b200 import _abi, _build
    hdr = open(os.path.join(ROOT, "include", "dmvae_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(dmvae_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    path = _build.build()
    lib = ctypes.CDLL(path)
    for name in declared:
        assert hasattr(lib, name), "libdmvae_b200.so does not export %s" % name
    assert declared == set(_abi.SIGNATURES.keys()), declared ^ set(_abi.SIGNATURES.keys())
    assert _abi.load().dmvae_abi_version() == 2


def test_abi_structs_match_c_layout():
    """sizeof of the ctypes mirrors must match what the C compiler lays out (checked by compiling a probe)."""
    import subprocess
    import tempfile
    from dmvae_b200 import _abi
    src = '#include <stdio.h>\n#include "dmvae_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(dmvae_gemm_epilogue), ' \
          'sizeof(dmvae_reparam_args), sizeof(dmvae_elbo_args), sizeof(dmvae_moe_args), sizeof(dmvae_chain_gemm));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "p.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "p.c"), "-o", os.path.join(d, "p")])
        sizes = [int(x) for x in subprocess.check_output([os.path.join(d, "p")]).split()]
    assert sizes == [ctypes.sizeof(_abi.GemmEpilogue), ctypes.sizeof(_abi.ReparamArgs), ctypes.sizeof(_abi.ElboArgs),
                     ctypes.sizeof(_abi.MoeArgs), ctypes.sizeof(_abi.ChainGemm)]


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from dmvae_b200.engine import Engine
    from dmvae_b200.session import Session
    with pytest.raises(RuntimeError):
        Session()
    with pytest.raises(RuntimeError):
        Engine(model="dmvae", input_type="binary", input_dim=784, latent_dim=10, n_classes=10, trunk=(500, 500), head=2000,
               decoder=(2000, 500, 500), name="dmvae")


def test_layout_matches_reference_parameter_counts_and_names():
    from dmvae_b200.engine import Layout, pad_dim
    from oracle import reference_graph as rg
    lay = Layout(model="dmvae", input_dim=784, latent_dim=10, n_classes=10, trunk=(500, 500), head=2000,
                 decoder=(2000, 500, 500), name="dmvae")
    assert lay.reference_parameter_count() == 4373014
    cfg = rg.GraphConfig()
    specs = {n: s for n, s, _ in rg.variable_specs(cfg)}
    names = set(lay.vars) | set(lay.dead_vars)
    assert names == set(specs)
    for n, vv in lay.vars.items():
        assert tuple(vv.shape) == tuple(specs[n]), n
    assert Layout(model="vade", input_dim=784, latent_dim=64, n_classes=50, trunk=(2000, 500, 500), head=0,
                  decoder=(500, 500, 2000), name="vade").reference_parameter_count() == 5745312
    assert Layout(model="dmvae", input_dim=3072, latent_dim=128, n_classes=100, trunk=(2000, 2000), head=4000,
                  decoder=(4000, 2000, 2000), name="dmvae").reference_parameter_count() == 46273028
    moe = Layout(model="dmvae", input_dim=784, latent_dim=1, n_classes=16, trunk=(500, 500), head=2000, decoder=(2000, 500, 500),
                 name="m", moe=dict(n_experts=16, output_dim=10, featLearn=False, scope="dmoe/dmoe/dmoe"))
    assert moe.reference_parameter_count() == 4330834 + 125600
    # padded layout: every input width leaves room for the ones column, all pitches are multiples of 64
    for ly in lay.layers.values():
        assert ly.in_pad > ly.n_in and ly.in_pad % 64 == 0 and ly.out_pad % 64 == 0 and ly.offset % 64 == 0
    assert pad_dim(784) == 832 and pad_dim(500) == 512 and pad_dim(2000) == 2048 and pad_dim(10) == 64 and pad_dim(63) == 64
    assert pad_dim(64) == 128
    with pytest.raises(NotImplementedError):
        Layout(model="gan", input_dim=4, latent_dim=2, n_classes=2, trunk=(4, 4), head=4, decoder=(4,), name="x")


def test_datasets_and_accuracy():
    from dmvae_b200.includes import utils
    rs = np.random.RandomState(0)
    X = (rs.uniform(size=(250, 12)) < .3).astype(np.float32)
    cls = rs.randint(0, 4, 250)
    ds = utils.Dataset((X, cls), batch_size=100)
    assert ds.epoch_len == 3 and len(ds) == 3
    b = list(ds.get_batches())
    assert [len(x) for x in b] == [100, 100, 50]                          # tail batch kept (utils.py:464-465)
    assert sorted(map(tuple, np.concatenate(b))) == sorted(map(tuple, X))
    assert ds.host_tensor().dtype.__str__() == "torch.uint8"              # {0,1} data is stored losslessly as uint8
    ds2 = utils.Dataset((X * 0.5, cls), batch_size=100, shuffle=False)
    assert ds2.host_tensor().dtype.__str__() == "torch.float32"
    assert np.array_equal(np.concatenate(list(ds2.get_batches())), X * 0.5)
    me = utils.MEDataset((X, cls, np.eye(4)[cls]), batch_size=64)
    xs, ys, cs = zip(*me.get_batches())
    assert sum(len(x) for x in xs) == 250 and ys[0].shape == (64, 4)
    # clustering accuracy: a permuted perfect clustering scores 1.0 (utils.py:22-34)
    perm = np.array([2, 0, 3, 1])
    w = np.eye(4)[perm[cls]]
    assert utils.get_clustering_accuracy(w, cls) == 1.0
    w[:25] = np.eye(4)[(perm[cls[:25]] + 1) % 4]
    assert abs(utils.get_clustering_accuracy(w, cls) - 0.9) < 1e-12
    g = utils.sample_gumbel((1000, 1, 10))
    assert g.shape == (1000, 1, 10) and abs(g.mean() - 0.5772) < 0.05
    with pytest.raises(NotImplementedError):
        utils.load_data("imagenet")
    sp = utils.load_data("spiral", classification=True)
    assert sp.train_data.shape == (25000, 2) and sp.input_type == "real" and sp.train_labels.shape == (25000, 5)
    sm = utils.load_data("synthetic_mnist", n_train=1000, n_test=200)
    assert sm.train_data.shape == (1000, 784) and set(np.unique(sm.train_data)) <= {0.0, 1.0}


def test_priors_keep_the_reference_assertions_and_samplers():
    from dmvae_b200 import priors
    nf = priors.NormalFactorial("z", 7)
    assert nf.sample_reparametrization_variable(5).shape == (5, 7)
    with pytest.raises(AssertionError):
        nf.kl_from_prior({"mean": None})
    with pytest.raises(AssertionError):
        nf.inverse_reparametrize(None, {"log_var": None})
    mix = priors.NormalMixtureFactorial("representation", 7, 3)
    with pytest.raises(AssertionError):
        mix.kl_from_prior({"mean": 0, "log_var": 0, "weights": 0})            # cluster_sample missing (priors.py:105-110)
    with pytest.raises(AssertionError):
        mix.sample_generative_feed(4)                                         # "session" missing (priors.py:71)
    df = priors.DiscreteFactorial("cluster", 1, 10)
    assert df.sample_reparametrization_variable(6).shape == (6, 1, 10)
    oh = df.sample_generative_feed(6)
    assert oh.shape == (6, 1, 10) and np.all(oh.sum(-1) == 1)
    with pytest.raises(AssertionError):
        df.inverse_reparametrize(None, {"logits": None})                      # temperature missing (priors.py:171)
    with pytest.raises(AssertionError):
        df.kl_from_prior({})
    for m in ("kl_from_prior", "sample_reparametrization_variable", "sample_generative_feed", "inverse_reparametrize"):
        with pytest.raises(NotImplementedError):
            getattr(priors.LatentVariable(), m)()


def test_model_api_surface_and_flags():
    from dmvae_b200 import base_models, models, nn, train
    m = base_models.DeepMixtureVAE("dmvae", "binary", 784, 10, 10, activation=nn.relu, initializer=nn.xavier_initializer)
    assert m.build_graph() is m
    for attr in ("X", "epsilon", "cluster", "mean", "log_var", "logits", "cluster_probs", "Z", "decoded_X", "reconstructed_X",
                 "reconstructed_Y_soft", "latent_variables", "kl_ratio", "is_training"):
        assert getattr(m, attr) is not None
    assert set(m.latent_variables) == {"C", "Z"}
    m.define_train_step(0.002, 100)
    assert m.train_step is not None and m.loss is not None and m.recon_loss is not None and m.latent_loss is not None
    feed = m.sample_reparametrization_variables(8)
    assert feed[m.epsilon].shape == (8, 10) and feed[m.cluster].shape == (8, 1, 10)
    m.define_pretrain_step(5e-4, 5e-4)
    assert m.vae_train_step is not None and m.prior_train_step is not None
    with pytest.raises(NotImplementedError):
        base_models.DeepMixtureVAE("x", "binary", 784, 10, 10, cnn=True)
    with pytest.raises(NotImplementedError):
        base_models.DeepMixtureVAE("x", "poisson", 784, 10, 10).build_graph()
    v = base_models.VaDE("vade", "real", 784, 10, 10, activation=nn.relu, initializer=nn.xavier_initializer).build_graph()
    assert v.latent_variables["C"][1] is None and "probs" in v.latent_variables["C"][2]
    assert list(v.sample_reparametrization_variables(4, variables=["Z"]).values())[0].shape == (4, 10)
    e = models.DeepMoE("dmoe", "binary", 784, 10, 16, True).build_graph()
    assert e.vae.latent_dim == 1 and e.lossVAE == 0 and e.n_classes == 16
    d = train.parser.parse_args([])
    ref = dict(model="dmvae", model_name="", dataset="mnist", latent_dim=10, output_dim=1, n_clusters=-1, n_experts=5,
               classification=False, n_epochs=500, pretrain_epochs_vae=200, pretrain_epochs_prior=200, init_lr=0.002,
               decay_rate=0.9, decay_epochs=25, pretrain=False, pretrain_vae_lr=0.0005, pretrain_decay_rate=0.9,
               pretrain_decay_epochs=25, pretrain_prior_lr=0.0005, kl_annealing=False, anneal_step=0.1, anneal_epochs=1000,
               plotting=False, plot_epochs=100, save_epochs=10, debug=False, visdom=False, featLearn=False)
    for k, val in ref.items():
        assert getattr(d, k) == val, k                                        # train.py:28-97


@pytest.mark.parametrize("fname,model", [("dmvae_small.npz", "dmvae"), ("vade_small.npz", "vade")])
def test_oracle_reproduces_golden_fixtures(fname, model):
    from oracle import reference_graph as rg
    z = np.load(os.path.join(ROOT, "tests", "golden", fname))
    if model == "dmvae":
        cfg = rg.GraphConfig(model="dmvae", input_dim=64, latent_dim=4, n_classes=5, trunk=(48, 40), head=56, decoder=(56, 40, 48))
    else:
        cfg = rg.GraphConfig.vade(input_dim=64, latent_dim=6, n_classes=7, trunk=(56, 40, 40), decoder=(40, 40, 56))
    V = {k[4:]: z[k] for k in z.files if k.startswith("var:")}
    out, g = rg.loss_and_grads(cfg, V, z["X"], z["eps"], kl_ratio=float(z["kl_ratio"]))
    for k in z.files:
        if k.startswith("out:"):
            assert np.allclose(out[k[4:]], z[k], rtol=1e-12, atol=1e-14), k
        if k.startswith("grad:"):
            assert np.allclose(g[k[5:]], z[k], rtol=1e-10, atol=1e-14), k
    # the fp32 twin of the oracle agrees with the fp64 truth within the fp32 tier tolerance
    import torch
    out32, _ = rg.loss_and_grads(cfg, V, z["X"], z["eps"], kl_ratio=float(z["kl_ratio"]), dtype=torch.float32)
    assert abs(out32["loss"] - z["out:loss"]) < 1e-5 * abs(z["out:loss"])


def test_dataset_storage_format_and_epoch_order():
    """Dataset keeps its rows in place and draws a permutation per epoch (utils.py:450-454); binarised data is stored as
    uint8 0/1, 8-bit intensities k/255 as uint8 with scale 1/255, anything else as float32."""
    from dmvae_b200.includes.utils import Dataset, MEDataset, _storage_format
    rs = np.random.RandomState(0)
    xb = (rs.uniform(size=(50, 12)) < 0.3).astype(np.float32)
    xi = (rs.randint(0, 256, size=(50, 12)) / 255.0).astype(np.float32)
    xr = rs.uniform(size=(50, 12)).astype(np.float32)
    assert _storage_format(xb) == (np.uint8, 1.0)
    assert _storage_format(xi) == (np.uint8, 1.0 / 255.0)
    assert _storage_format(xr) == (np.float32, 1.0)
    d = Dataset((xi, np.arange(50)), batch_size=16)
    h = d.host_tensor()
    assert h.dtype.__str__() == "torch.uint8" and d.host_scale == 1.0 / 255.0
    assert np.array_equal(h.numpy(), np.rint(xi * 255).astype(np.uint8))
    np.random.seed(1)
    batches = list(d.get_batches())
    assert sorted(d.perm.tolist()) == list(range(50)) and len(batches) == 4 and len(batches[-1]) == 2
    assert np.array_equal(np.concatenate(batches), xi[d.perm]) and np.array_equal(d.epoch_classes(), d.perm)
    d.prefetch_epoch()
    nxt = d._next_perm.copy()
    d.begin_epoch()
    assert np.array_equal(d.perm, nxt) and d._next_perm is None
    me = MEDataset((xb, np.arange(50), np.eye(5)[np.arange(50) % 5]), batch_size=20)
    got = list(me.get_batches())
    assert np.array_equal(np.concatenate([g[2] for g in got]), me.perm) and got[0][1].shape == (20, 5)


def test_pack_bits_layout():
    """includes/utils.py::_pack_bits: little bit order, rows padded to 16 bytes, None for non-binary data."""
    from dmvae_b200.includes.utils import _pack_bits
    rs = np.random.RandomState(0)
    X = (rs.uniform(size=(5, 784)) < 0.5).astype(np.float32)
    p = _pack_bits(X)
    assert p.shape == (5, 112) and p.dtype == np.uint8 and not p[:, 98:].any()
    back = np.unpackbits(p[:, :98], axis=1, bitorder="little")[:, :784]
    assert np.array_equal(back, X.astype(np.uint8))
    assert _pack_bits(rs.uniform(size=(4, 32)).astype(np.float32)) is None
    assert _pack_bits(np.zeros((4, 20), np.float32)) is None          # width not a multiple of 16
