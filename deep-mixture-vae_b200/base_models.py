"""VAE base class, DeepMixtureVAE and VaDE with the API surface of code/base_models.py.

The reference builds a TensorFlow graph; here ``build_graph`` records the layer shapes and creates the named
handles (``X``, ``epsilon``, ``mean``, ``logits``, ``loss``, ``train_step`` ...), and the first use with a Session
creates the step engine (flat parameter buffers + padded activation buffers on the device).  ``session.run``
on those handles, ``train_op``, ``get_accuracy`` and the pre-training entry points keep the reference's names,
argument order and return values.  Only the MLP encoder (cnn=False, base_models.py:218-226) is implemented.
"""
from __future__ import annotations

import os
from typing import Any, Dict, List, Optional

import numpy as np
import torch
from tqdm import tqdm

from . import nn, priors
from .engine import AdamState, Engine
from .includes.network import DeepNetwork
from .includes.utils import Dataset, get_clustering_accuracy, hungarian_accuracy
from .session import Handle, Session
from . import _abi

import ctypes as C


def _to_device(a, device, dtype=None) -> torch.Tensor:
    if isinstance(a, torch.Tensor):
        t = a
    else:
        t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.to(device, non_blocking=True)


class VAE:
    def __init__(self, name, input_type, input_dim, latent_dim, activation=None, initializer=None):
        self.name = name
        self.input_dim = input_dim
        self.latent_dim = latent_dim
        self.input_type = input_type
        self.activation = activation
        self.initializer = initializer
        nn.check_activation(activation)
        nn.check_initializer(initializer)
        self.path = ""
        self.kl_ratio = Handle("kl_ratio", self, "placeholder")          # placeholder_with_default(1.0)
        self.is_training = Handle("is_training", self, "placeholder")    # placeholder_with_default(True)
        self.X = None
        self.decoded_X = None
        self.train_step = None
        self.latent_variables = dict()
        self.engine: Optional[Engine] = None
        self._lr = {}
        self._pending_vars: Dict[str, np.ndarray] = {}
        # engine options (new; defaults reproduce the reference maths)
        self.gemm_dtype = "bf16"
        self.seed = 0
        self.max_batch = 4096
        self.cluster_sample = False
        self.temperature = 1.0

    # -- graph ----------------------------------------------------------------------------------
    def build_graph(self, encoder_layer_sizes=None, decoder_layer_sizes=None):
        raise NotImplementedError

    def _make_handles(self, names_tensors, names_ops=()):
        for n in names_tensors:
            setattr(self, n, Handle(n, self, "tensor"))
        for n in names_ops:
            setattr(self, n, Handle(n, self, "op"))

    def _engine_kwargs(self) -> dict:
        raise NotImplementedError

    def _ensure_engine(self, session: Optional[Session] = None) -> Engine:
        if self.engine is None:
            kw = self._engine_kwargs()
            if session is not None and session.gemm_dtype is not None:
                kw["gemm_dtype"] = session.gemm_dtype
            kw["device"] = session.device if session is not None else None
            self.engine = Engine(**kw)
            for lv, _, _ in self.latent_variables.values():
                if hasattr(lv, "attach"):
                    lv.attach(self.engine)
            if self._pending_vars:
                self.engine.load_variables(self._pending_vars)
                self._pending_vars = {}
        return self.engine

    # -- variables (checkpoint interchange under the reference's TF variable names) ---------------
    def get_variables(self) -> Dict[str, np.ndarray]:
        return self._ensure_engine().state_dict()

    def set_variables(self, values: Dict[str, np.ndarray]):
        if self.engine is None:
            self._pending_vars.update(values)
        else:
            self.engine.load_variables(values)

    def save(self, ckpt_path: str):
        """Replaces tf.train.Saver(TRAINABLE_VARIABLES).save: one .npz keyed by the TF variable names
        (Adam slots are not saved, exactly as in the reference: train.py:233-236)."""
        d = os.path.dirname(ckpt_path)
        if d:
            os.makedirs(d, exist_ok=True)
        np.savez(ckpt_path if ckpt_path.endswith(".npz") else ckpt_path + ".npz", **self.get_variables())

    def restore(self, ckpt_path: str):
        p = ckpt_path if ckpt_path.endswith(".npz") else ckpt_path + ".npz"
        with np.load(p) as z:
            self.set_variables({k: z[k] for k in z.files})

    # -- noise ----------------------------------------------------------------------------------
    def sample_reparametrization_variables(self, n, variables=None):
        """base_models.py:44-56: {epsilon handle: host noise} for every latent variable that has one."""
        samples = dict()
        if variables is None:
            for lv, eps, _ in self.latent_variables.values():
                if eps is not None:
                    samples[eps] = lv.sample_reparametrization_variable(n)
        else:
            for var in variables:
                lv, eps, _ = self.latent_variables[var]
                if eps is not None:
                    samples[eps] = lv.sample_reparametrization_variable(n)
        return samples

    def sample_generative_feed(self, n, **kwargs):
        samples = dict()
        for name, (lv, _, _) in self.latent_variables.items():
            kwargs_ = dict() if name not in kwargs else kwargs[name]
            samples[name] = lv.sample_generative_feed(n, **kwargs_)
        return samples

    # -- losses / train step --------------------------------------------------------------------
    def define_latent_loss(self):
        self.latent_loss = Handle("latent_loss", self, "tensor")

    def define_recon_loss(self):
        if self.input_type not in ("binary", "real"):
            raise NotImplementedError
        self.recon_loss = Handle("recon_loss", self, "tensor")

    def define_train_loss(self):
        self.define_latent_loss()
        self.define_recon_loss()
        self.loss = Handle("loss", self, "tensor")

    def define_train_step(self, init_lr, decay_steps, decay_rate=0.9):
        """base_models.py:95-110.  The reference passes the literal ``global_step=0`` to exponential_decay, so the
        learning rate is the constant ``init_lr``; decay_steps / decay_rate are accepted and inert, as there."""
        self._lr["train"] = float(init_lr)
        self.define_train_loss()
        self.train_step = Handle("train_step", self, "op")

    # -- execution ------------------------------------------------------------------------------
    _TRAIN_OPS = {"train_step": ("train", "all"), "vae_train_step": ("vae", "vae"), "prior_train_step": ("prior", "prior")}

    def _run(self, session: Session, fetches: List[Handle], feed: Dict[Handle, Any]):
        eng = self._ensure_engine(session)
        names = [f.name for f in fetches]
        fd = {k.name: v for k, v in feed.items() if k.owner is self}
        if "X" not in fd and "Z" in fd:
            # generation: the latent code is fed directly (includes/visualization.py:83-87) - decoder only
            bad = [n for n in names if n not in ("decoded_X", "reconstructed_X", "Z")]
            if bad:
                raise ValueError("with Z fed only decoded_X / reconstructed_X can be fetched (got %r)" % bad)
            Zd = _to_device(fd["Z"], eng.device, torch.float32).contiguous()
            rows = len(Zd)
            eng.decode_latent(Zd, rows)
            return [self._fetch(eng, n, rows) for n in names]
        if "X" not in fd:
            raise ValueError("feed_dict must contain the X placeholder")
        X = fd["X"]
        rows = len(X)
        Xd = _to_device(X, eng.device, torch.float32 if not (isinstance(X, torch.Tensor) and X.dtype == torch.uint8) else None)
        eps = fd.get("epsilon_Z", fd.get("epsilon"))
        eps_d = _to_device(eps, eng.device, torch.float32) if eps is not None else None
        gum = fd.get("epsilon_C")
        gum_d = _to_device(gum, eng.device, torch.float32) if (gum is not None and eng.cluster_sample) else None
        kl_ratio = float(fd.get("kl_ratio", 1.0))
        train = [n for n in names if n in self._TRAIN_OPS]
        if train:
            key, mode = self._TRAIN_OPS[train[0]]
            opt = eng.optimizer(key, self._lr[key])
            klr, rs = kl_ratio, 1.0
            if mode == "vae":
                klr = 0.0                 # minimize(recon_loss): the KL terms contribute no gradient
            elif mode == "prior":
                rs = 0.0                  # minimize(latent_loss, var_list = c-head)
                klr = 1.0
            # session.run feeds are transient device tensors: a captured graph keyed on their address would be rebuilt for
            # every new allocation, so only the epoch loop's persistent staging buffers are captured (Engine.run_epoch)
            eng.train_step(Xd, rows, opt, eps_d, gum_d, klr, mode, rs, graph=False)
        else:
            need_full = any(n not in ("logits", "mean", "log_var") for n in names) or eng.model == "vade"
            if need_full:
                eng.forward_backward(Xd, rows, eps_d, gum_d, kl_ratio, backward=False)
            else:
                eng.stage_input(Xd, rows)
                eng.encode(rows)
        return [self._fetch(eng, n, rows) for n in names]

    def _fetch(self, eng: Engine, name: str, rows: int):
        L, K, D = eng.L, eng.K, eng.D
        if name in self._TRAIN_OPS:
            return None
        if name in ("loss", "recon_loss", "latent_loss", "vae_loss"):
            lo = eng.loss_out.cpu().numpy()
            return {"loss": float(lo[3]), "recon_loss": float(lo[0]), "vae_loss": float(lo[0]),
                    "latent_loss": float(lo[1] + lo[2])}[name]
        t = {"mean": lambda: eng.zh[:rows, :L], "log_var": lambda: eng.zh[:rows, L:2 * L],
             "logits": lambda: eng.ch[:rows, :K], "cluster_probs": lambda: eng.qc[:rows],
             "Z": lambda: eng.zb[:rows, :L], "decoded_X": lambda: eng.decoded[:rows, :D],
             # base_models.py:295-300: sigmoid(decoded_X) for binary inputs - the output layer's GEMM with the sigmoid in
             # its epilogue (DMVAE_ACT_SIGMOID), computed on demand
             "reconstructed_X": lambda: eng.reconstruct(rows)[:rows, :D]}.get(name)
        if t is None:
            raise KeyError("cannot fetch %r" % name)
        return t().float().cpu().numpy()

    def decode(self, session, Z):
        """Decoder only: reconstructed_X for given latent codes (what includes/visualization.py:83-87 does by feeding
        ``model.Z``)."""
        return session.run(self.reconstructed_X, feed_dict={self.Z: Z})

    def reconstruct(self, session, X):
        """reconstructed_X with zero noise (includes/visualization.py:39-46)."""
        eps = self.epsilon if getattr(self, "epsilon", None) is not None else None
        feed = {self.X: X}
        if eps is not None:
            feed[eps] = np.zeros((len(X), self.latent_dim), np.float32)
        return session.run(self.reconstructed_X, feed_dict=feed)

    def train_op(self, session, data, kl_ratio=1.0):
        """base_models.py:112-132: one epoch; returns the mean batch loss.

        With this package's ``Dataset`` the epoch runs device-side: batches are contiguous slices of a pinned
        host array copied asynchronously (double-buffered) while the previous step computes, the noise comes from
        the device Philox generator and the loss is accumulated without a per-step synchronisation.  Any other
        object with ``get_batches()`` goes through the reference-shaped loop below."""
        assert(self.train_step is not None)
        if isinstance(data, Dataset):
            return self._train_epoch_fast(session, data, kl_ratio)
        loss = 0.0
        for batch in data.get_batches():
            feed = {self.X: batch, self.is_training: True, self.kl_ratio: kl_ratio}
            feed.update(self.sample_reparametrization_variables(len(batch)))
            batch_loss, _ = session.run([self.loss, self.train_step], feed_dict=feed)
            loss += batch_loss / data.epoch_len
        return loss

    def _train_epoch_fast(self, session, data: Dataset, kl_ratio=1.0, op="train_step"):
        eng = self._ensure_engine(session)
        key, mode = self._TRAIN_OPS[op]
        opt = eng.optimizer(key, self._lr[key])
        data.begin_epoch()                                     # draws the epoch's permutation (utils.py:450-454)
        runner = eng.dp if eng.dp is not None else eng
        bits = data.host_bits()
        if bits is not None:                                   # binarised rows: one bit per element over the bus
            return runner.run_epoch(bits[0], data.batch_size, opt, kl_ratio, mode, perm=data.perm,
                                    while_busy=data.prefetch_epoch, x_scale=1.0, packed_D=bits[1])
        host = data.host_tensor()
        return runner.run_epoch(host, data.batch_size, opt, kl_ratio, mode, perm=data.perm,
                                while_busy=data.prefetch_epoch, x_scale=data.host_scale)

    def debug(self, session, data):
        """base_models.py:134-147 drops into pdb; here the prepared feed is returned instead."""
        for batch in data.get_batches():
            feed = {self.X: batch}
            feed.update(self.sample_reparametrization_variables(len(batch)))
            return feed


class DeepMixtureVAE(VAE):
    """base_models.py:150-432 (MLP encoder branch)."""

    def __init__(self, name, input_type, input_dim, latent_dim, n_classes, activation=None, initializer=None, cnn=False,
                 hidden=(500, 500, 2000), decoder=(2000, 500, 500)):
        VAE.__init__(self, name, input_type, input_dim, latent_dim, activation=activation, initializer=initializer)
        self.n_classes = n_classes
        if cnn:
            raise NotImplementedError("the CNN encoder (base_models.py:178-216) is outside the accelerated path; "
                                      "use cnn=False (the MLP branch, base_models.py:218-226)")
        self.cnn = False
        self.hidden = tuple(hidden)          # (trunk 1, trunk 2, head width); reference: 500, 500, 2000
        self.decoder_sizes = tuple(decoder)  # reference: 2000, 500, 500

    def build_graph(self):
        if self.input_type not in ("binary", "real"):
            raise NotImplementedError
        self.X = Handle("X", self, "placeholder")
        self.epsilon = Handle("epsilon_Z", self, "placeholder")
        self.cluster = Handle("epsilon_C", self, "placeholder")
        self._make_handles(["mean", "log_var", "logits", "cluster_probs", "Z", "decoded_X", "reconstructed_X",
                            "reconstructed_Y_soft"])
        self.latent_variables = dict()
        self.latent_variables.update({
            "C": (priors.DiscreteFactorial("cluster", 1, self.n_classes), self.cluster, {"logits": self.logits}),
            "Z": (priors.NormalMixtureFactorial("representation", self.latent_dim, self.n_classes), self.epsilon,
                  {"mean": self.mean, "log_var": self.log_var, "weights": self.cluster_probs,
                   "cluster_sample": self.cluster_sample}),
        })
        self.decoder_network = DeepNetwork(
            "layers", [("fc", {"input_dim": a, "output_dim": b})
                       for a, b in zip((self.latent_dim,) + self.decoder_sizes[:-1], self.decoder_sizes)],
            activation=self.activation or nn.relu, initializer=self.initializer or nn.xavier_initializer)
        return self

    def _engine_kwargs(self):
        return dict(model="dmvae", input_type=self.input_type, input_dim=self.input_dim, latent_dim=self.latent_dim,
                    n_classes=self.n_classes, trunk=self.hidden[:2], head=self.hidden[2], decoder=self.decoder_sizes,
                    name=self.name, gemm_dtype=self.gemm_dtype, seed=self.seed, max_rows=self.max_batch,
                    cluster_sample=self.cluster_sample, temperature=self.temperature, moe=getattr(self, "moe_config", None))

    def define_pretrain_step(self, vae_lr, prior_lr):
        """base_models.py:304-321: two more Adam instances - recon_loss over everything, latent_loss over the
        c-head variables only."""
        self.define_train_loss()
        self.vae_loss = Handle("vae_loss", self, "tensor")
        self._lr["vae"] = float(vae_lr)
        self.vae_train_step = Handle("vae_train_step", self, "op")
        self._lr["prior"] = float(prior_lr)
        self.prior_train_step = Handle("prior_train_step", self, "op")

    def _pretrain_loop(self, session, data, n_epochs, loss_handle, step_handle, ckpt_path):
        min_loss = float("inf")
        with tqdm(range(n_epochs), disable=n_epochs == 0) as bar:
            for _ in bar:
                loss = 0
                for batch in data.get_batches():
                    feed = {self.X: batch, self.epsilon: np.zeros((len(batch), self.latent_dim)), self.is_training: True}
                    batch_loss, _ = session.run([loss_handle, step_handle], feed_dict=feed)
                    loss += batch_loss / data.epoch_len
                bar.set_postfix({"loss": "%.4f" % loss})
                if loss <= min_loss:
                    min_loss = loss
                    self.save(ckpt_path)

    def pretrain_vae(self, session, data, n_epochs):
        ckpt_path = self.path + "/vae/parameters.ckpt"
        try:
            self.restore(ckpt_path)
        except Exception:
            print("Could not load trained ae parameters")
        self._pretrain_loop(session, data, n_epochs, self.recon_loss, self.vae_train_step, ckpt_path)

    def _gmm_init(self, session, data, n_epochs, n_init, ckpt_path):
        """base_models.py:366-389: fit a diagonal GMM on the posterior means and load it into the prior tables."""
        from sklearn.mixture import GaussianMixture
        Z = []
        for i in range(0, len(data.data), 4096):
            Z.append(session.run(self.mean, feed_dict={self.X: data.data[i:i + 4096]}))
        Z = np.concatenate(Z, axis=0)
        gmm_model = GaussianMixture(n_components=self.n_classes, covariance_type="diag", max_iter=n_epochs, n_init=n_init,
                                    weights_init=np.ones(self.n_classes) / self.n_classes)
        gmm_model.fit(Z)
        self.set_variables({self.name + "/representation/means": gmm_model.means_,
                            self.name + "/representation/log_vars": np.log(gmm_model.covariances_ + 1e-20)})
        self.save(ckpt_path)

    def pretrain_prior(self, session, data, n_epochs):
        ckpt_path = self.path + "/prior/parameters.ckpt"
        try:
            self.restore(ckpt_path)
        except Exception:
            print("Could not load trained prior parameters")
            if n_epochs > 0:
                self._gmm_init(session, data, n_epochs, 20, ckpt_path)
        self._pretrain_loop(session, data, n_epochs, self.latent_loss, self.prior_train_step, ckpt_path)

    def pretrain(self, session, data, n_epochs_vae, n_epochs_gmm):
        assert(self.vae_train_step is not None and self.prior_train_step is not None)
        self.pretrain_vae(session, data, n_epochs_vae)
        self.pretrain_prior(session, data, n_epochs_gmm)

    def get_accuracy(self, session, data):
        """base_models.py:425-432: encoder + c-head only, argmax over the logits, Hungarian matching.
        The argmax and the contingency counts are produced on the device; only the K x K matrix comes back."""
        eng = self._ensure_engine(session)
        K = self.n_classes
        n_labels = max(K, int(np.max(data.classes)) + 1)
        counts = torch.zeros(K, n_labels, dtype=torch.int32, device=eng.device)
        total = 0
        bs = max(1, min(eng.max_rows, 4096))                  # chunked to the buffers the engine already has
        classes = torch.from_numpy(np.asarray(data.classes).astype(np.int32)).to(eng.device)
        # batches come from the dataset's pinned host copy (storage dtype: 1 byte / pixel for binarised data), staged
        # asynchronously into a persistent device buffer; rows stay in their original order, like data.classes
        host = data.host_tensor() if isinstance(data, Dataset) else None
        if host is not None:
            eng.x_scale = data.host_scale
        buf = None
        if host is not None:
            buf = getattr(eng, "_eval_stage", None)
            if buf is None or buf.shape[0] < bs or buf.dtype != host.dtype:
                buf = eng._eval_stage = torch.empty(bs, eng.D, dtype=host.dtype, device=eng.device)
        for i in range(0, len(data.data), bs):
            rows = min(bs, len(data.data) - i)
            if host is not None:
                buf[:rows].copy_(host[i:i + rows], non_blocking=True)
                Xd = buf
            else:
                Xd = _to_device(data.data[i:i + rows], eng.device, torch.float32)
            eng.stage_input(Xd, rows)
            eng.encode(rows, heads=("c",))
            _abi.check(eng.lib.dmvae_argmax_contingency(eng.ctx, eng.ch.data_ptr(), eng.ch.stride(0), rows, K,
                                                        classes[i:i + rows].data_ptr(), n_labels, eng.argmax.data_ptr(),
                                                        counts.data_ptr(), eng._stream()))
            total += rows
        d = counts.cpu().numpy().astype(np.int64)
        if n_labels > K:
            d = np.pad(d, ((0, n_labels - K), (0, 0)))
        elif n_labels < K:
            d = np.pad(d, ((0, 0), (0, K - n_labels)))
        return hungarian_accuracy(d, total)


class VaDE(VAE):
    """base_models.py:435-670 (MLP encoder D-2000-500-500, decoder L-500-500-2000-D)."""

    def __init__(self, name, input_type, input_dim, latent_dim, n_classes, activation=None, initializer=None, cnn=False,
                 hidden=(2000, 500, 500), decoder=(500, 500, 2000)):
        VAE.__init__(self, name, input_type, input_dim, latent_dim, activation=activation, initializer=initializer)
        self.n_classes = n_classes
        if cnn:
            raise NotImplementedError("the CNN encoder is outside the accelerated path; use cnn=False")
        self.cnn = False
        self.hidden = tuple(hidden)
        self.decoder_sizes = tuple(decoder)

    def build_graph(self):
        if self.input_type not in ("binary", "real"):
            raise NotImplementedError
        self.X = Handle("X", self, "placeholder")
        self.epsilon = Handle("epsilon", self, "placeholder")
        self._make_handles(["mean", "log_var", "cluster_probs", "Z", "decoded_X", "reconstructed_X"])
        self.latent_variables = dict()
        params = {"mean": self.mean, "log_var": self.log_var, "cluster_sample": False}
        self.latent_variables.update({
            "Z": (priors.NormalMixtureFactorial("representation", self.latent_dim, self.n_classes), self.epsilon, params)
        })
        params["weights"] = self.cluster_probs
        self.latent_variables.update({
            "C": (priors.DiscreteFactorial("cluster", 1, self.n_classes), None, {"probs": self.cluster_probs})
        })
        self.encoder_network = DeepNetwork(
            "layers", [("fc", {"input_dim": a, "output_dim": b})
                       for a, b in zip((self.input_dim,) + self.hidden[:-1], self.hidden)],
            activation=self.activation or nn.relu, initializer=self.initializer or nn.xavier_initializer)
        self.decoder_network = DeepNetwork(
            "layers", [("fc", {"input_dim": a, "output_dim": b})
                       for a, b in zip((self.latent_dim,) + self.decoder_sizes[:-1], self.decoder_sizes)],
            activation=self.activation or nn.relu, initializer=self.initializer or nn.xavier_initializer)
        return self

    def _engine_kwargs(self):
        return dict(model="vade", input_type=self.input_type, input_dim=self.input_dim, latent_dim=self.latent_dim,
                    n_classes=self.n_classes, trunk=self.hidden, head=0, decoder=self.decoder_sizes, name=self.name,
                    gemm_dtype=self.gemm_dtype, seed=self.seed, max_rows=self.max_batch, moe=getattr(self, "moe_config", None))

    def define_pretrain_step(self, vae_lr, _prior_lr=None):
        """base_models.py:574-580."""
        self.define_train_loss()
        self.vae_loss = Handle("vae_loss", self, "tensor")
        self._lr["vae"] = float(vae_lr)
        self.vae_train_step = Handle("vae_train_step", self, "op")

    pretrain_vae = DeepMixtureVAE.pretrain_vae
    _pretrain_loop = DeepMixtureVAE._pretrain_loop
    _gmm_init = DeepMixtureVAE._gmm_init

    def pretrain_prior(self, session, data, n_epochs):
        """base_models.py:611-652: GMM initialisation only (no gradient stage for VaDE)."""
        ckpt_path = self.path + "/prior/parameters.ckpt"
        try:
            self.restore(ckpt_path)
        except Exception:
            print("Could not load pretrained prior parameters")
            if n_epochs > 0:
                self._gmm_init(session, data, n_epochs, 5, ckpt_path)

    def pretrain(self, session, data, n_epochs_vae, n_epochs_prior):
        assert(self.vae_train_step is not None)
        self.pretrain_vae(session, data, n_epochs_vae)
        self.pretrain_prior(session, data, n_epochs_prior)

    def get_accuracy(self, session, data, k=10):
        """base_models.py:654-670: mean of q(c|z) over k noise draws, then Hungarian matching."""
        weights = None
        bs = 4096
        for _ in range(k):
            parts = []
            for i in range(0, len(data.data), bs):
                X = data.data[i:i + bs]
                feed = {self.X: X}
                feed.update(self.sample_reparametrization_variables(len(X), variables=["Z"]))
                parts.append(session.run(self.cluster_probs, feed_dict=feed))
            w = np.concatenate(parts, axis=0)
            weights = w if weights is None else weights + w
        weights = weights / k
        return get_clustering_accuracy(weights, data.classes)
