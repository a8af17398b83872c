"""Mixture-of-experts wrappers with the API of code/models.py: the gate is the cluster posterior of a clustering VAE
(models.py:74), the experts are linear / softmax regressors evaluated by ONE dense GEMM over all experts (the
reference tiles the input n_experts times, models.py:76-81)."""
from __future__ import annotations

import numpy as np
import torch
from tqdm import tqdm

from . import nn
from .base_models import DeepMixtureVAE, VaDE, _to_device
from .includes.utils import Dataset, MEDataset, get_clustering_accuracy
from .session import Handle


class MoE:
    is_moe = True

    def __init__(self, name, input_type, input_dim, latent_dim, output_dim, n_experts, classification, activation=None,
                 initializer=None, lossVAE=1, featLearn=1, cnn=1):
        self.name = name
        self.input_dim = input_dim
        self.latent_dim = latent_dim
        self.output_dim = output_dim
        self.input_type = input_type
        self.classification = classification
        self.n_experts = self.n_classes = n_experts
        self.activation = activation
        self.initializer = initializer
        self.vae = None
        self.featLearn = featLearn
        self.lossVAE = lossVAE
        self.cnn = False          # only the MLP encoder is accelerated (SURVEY note A); the flag is accepted
        self.train_step = None
        self._lr = {}
        self.gemm_dtype = "bf16"

    def _define_vae(self):
        raise NotImplementedError

    def define_vae(self):
        self._define_vae()

    def build_graph(self):
        self.define_vae()
        self.vae.moe_config = dict(n_experts=self.n_experts, output_dim=self.output_dim, featLearn=bool(self.featLearn),
                                   lossVAE=bool(self.lossVAE), classification=bool(self.classification),
                                   scope="/".join([self.name] * 3))      # three nested scopes, models.py:36-41
        self.X = self.vae.X
        self.Z = self.vae.Z
        self.Y = Handle("Y", self, "placeholder")
        self.reconstructed_X = self.vae.reconstructed_X
        self.expert_probs = self.vae.cluster_probs
        for n in ("reconstructed_Y_soft", "reconstructed_Y", "error"):
            setattr(self, n, Handle(n, self, "tensor"))
        self.regression_weights = Handle("regression_weights", self, "tensor")
        self.regression_biases = Handle("regression_biases", self, "tensor")
        return self

    def sample_generative_feed(self, n, **kwargs):
        return self.vae.sample_generative_feed(n, **kwargs)

    def sample_reparametrization_variables(self, n):
        return self.vae.sample_reparametrization_variables(n)

    def define_train_loss(self):
        self.vae.define_train_loss()
        self.recon_loss = Handle("moe_recon_loss", self, "tensor")
        self.loss = Handle("moe_loss", self, "tensor")

    def define_pretrain_step(self, init_lr, decay_steps, decay_rate=0.9):
        self.vae.define_train_step(init_lr, decay_steps, decay_rate)

    def define_train_step(self, init_lr, decay_steps, decay_rate=0.9, pretrain_init_lr=None, pretrain_decay_steps=None,
                          pretrain_decay_rate=None):
        """models.py:170-183: constant learning rate (global_step=0), one Adam over everything with a gradient."""
        self.define_train_loss()
        self._lr["moe"] = float(init_lr)
        self.train_step = Handle("moe_train_step", self, "op")

    # -- execution --------------------------------------------------------------------------------
    def _engine(self, session):
        self.vae.gemm_dtype = self.gemm_dtype
        return self.vae._ensure_engine(session)

    def _run(self, session, fetches, feed):
        eng = self._engine(session)
        fd = {}
        for k, v in feed.items():
            fd[k.name] = v
        X, Y = fd["X"], fd["Y"]
        rows = len(X)
        Xd = _to_device(X, eng.device, torch.float32)
        Yd = _to_device(Y, eng.device, torch.float32)
        eps = fd.get("epsilon_Z", fd.get("epsilon"))
        eps_d = _to_device(eps, eng.device, torch.float32) if eps is not None else None
        names = [f.name for f in fetches]
        train = "moe_train_step" in names
        opt = eng.optimizer("moe", self._lr["moe"]) if train else None
        eng.moe_step(Xd, Yd, rows, opt, eps_d, None, float(fd.get("kl_ratio", 1.0)), train=train)
        ml = eng.moe_loss.cpu().numpy()
        lo = eng.loss_out.cpu().numpy()
        sup = float(ml[0]) / rows                              # batch mean of the supervised loss summands
        err = float(ml[1]) if self.classification else float(ml[1]) / rows
        out = []
        for f in fetches:
            n = f.name
            if n == "moe_train_step":
                out.append(None)
            elif n == "moe_recon_loss":
                out.append(sup)
            elif n == "moe_loss":
                out.append(sup + (float(lo[3]) if self.lossVAE else 0.0))
            elif n == "error":
                out.append(err)
            elif n in ("reconstructed_Y_soft", "reconstructed_Y"):
                ys = eng.moe_ysoft[:rows].cpu().numpy()
                if n == "reconstructed_Y" and self.classification:
                    ys = np.eye(self.output_dim)[eng.moe_cls[:rows].cpu().numpy()]
                out.append(ys)
            elif n == "regression_weights":
                out.append(eng.get_variable(self.vae.moe_config["scope"] + "/regression_weights"))
            elif n == "regression_biases":
                out.append(eng.get_variable(self.vae.moe_config["scope"] + "/regression_biases"))
            else:
                out.append(self.vae._fetch(eng, n, rows))
        return out

    def get_accuracy(self, session, data):
        """models.py:121-147.  Two notes: (1) the reference fetches ``self.vae.logits``, which its VaDE does not define
        (VaDEMoE.get_accuracy raises AttributeError there); here a VaDE gate is scored by its cluster posterior, whose
        argmax is what the clustering accuracy needs.  (2) the batches come in the epoch's shuffled order, so the
        classes are collected from the batches themselves (the reference's MEDataset re-orders ``data.classes`` in place)."""
        error = 0.0
        logits, classes = [], []
        gate = self.vae.logits if hasattr(self.vae, "logits") else self.vae.cluster_probs
        for X_batch, Y_batch, C_batch in data.get_batches():
            feed = {self.X: X_batch, self.Y: Y_batch}
            feed.update(self.vae.sample_reparametrization_variables(len(X_batch)))
            batchLogits, batchError = session.run([gate, self.error], feed_dict=feed)
            error += batchError
            logits.append(batchLogits)
            classes.append(C_batch)
        logits = np.concatenate(logits, axis=0)
        accClustering = get_clustering_accuracy(logits, np.concatenate(classes, axis=0))
        if self.classification:
            error /= data.len
            return 1 - error, accClustering
        error /= data.epoch_len
        return -error, accClustering

    def pretrain(self, session, data, n_epochs):
        print("Pretraining Model")
        data = Dataset((data.data, data.classes), data.batch_size, data.shuffle)
        with tqdm(range(n_epochs)) as bar:
            for _ in bar:
                self.vae.train_op(session, data)

    def train_op(self, session, data, kl_ratio=1.0):
        """models.py:194-221: returns (loss, batch_acc of the last batch, lossCls)."""
        assert(self.train_step is not None)
        if isinstance(data, MEDataset):
            return self._train_epoch_fast(session, data, kl_ratio)
        loss = 0.0
        lossCls = 0.0
        batch_error, Y_batch = 0.0, None
        for X_batch, Y_batch, _ in data.get_batches():
            feed = {self.X: X_batch, self.Y: Y_batch, self.vae.kl_ratio: kl_ratio}
            feed.update(self.vae.sample_reparametrization_variables(len(X_batch)))
            batch_error, batch_loss, _, batch_lossCls = session.run(
                [self.error, self.loss, self.train_step, self.recon_loss], feed_dict=feed)
            lossCls += batch_lossCls / data.epoch_len
            loss += batch_loss / data.epoch_len
        if self.classification:
            batch_acc = 1 - batch_error / Y_batch.shape[0]
        else:
            batch_acc = -batch_error
        return loss, batch_acc, lossCls

    def _train_epoch_fast(self, session, data, kl_ratio=1.0):
        """models.py:194-221 with this package's MEDataset: the epoch's permutation goes to the device, which gathers each
        batch (X and Y) out of the pinned host copies; noise comes from the device Philox generator; the step is a replayed
        CUDA graph; one synchronisation per epoch.  Same return value: (loss, batch_acc of the last batch, lossCls)."""
        eng = self._engine(session)
        opt = eng.optimizer("moe", self._lr["moe"])
        data.begin_epoch()
        hx, hy = data.host_tensors()
        bits = data.host_bits()
        if bits is not None:                                    # binarised inputs: one bit per element over the bus
            log = eng.run_epoch_moe(bits[0], hy, data.batch_size, opt, kl_ratio, perm=data.perm,
                                    while_busy=data.prefetch_epoch, x_scale=1.0, packed_D=bits[1])
        else:
            log = eng.run_epoch_moe(hx, hy, data.batch_size, opt, kl_ratio, perm=data.perm, while_busy=data.prefetch_epoch,
                                    x_scale=data.host_scale)
        nb = len(log)
        last_rows = data.len - (nb - 1) * data.batch_size if nb == data.epoch_len else data.batch_size
        rows = np.full(nb, data.batch_size, np.float64)
        rows[-1] = last_rows
        sup = log[:, 0] / rows                                  # batch means of the supervised loss
        lossCls = float(np.sum(sup) / data.epoch_len)
        loss = float(np.sum(sup + (log[:, 5] if self.lossVAE else 0.0)) / data.epoch_len)
        if self.classification:
            batch_acc = 1 - float(log[-1, 1]) / last_rows
        else:
            batch_acc = -float(log[-1, 1]) / last_rows
        return loss, batch_acc, lossCls

    def debug(self, session, data, kl_ratio=1.0):
        for X_batch, Y_batch, _ in data.get_batches():
            feed = {self.X: X_batch, self.Y: Y_batch, self.vae.kl_ratio: kl_ratio}
            feed.update(self.vae.sample_reparametrization_variables(len(X_batch)))
            return feed


class DeepMoE(MoE):
    """models.py:240-250: lossVAE=0, latent_dim=1 - only the supervised loss trains the gate and the experts."""

    def __init__(self, name, input_type, input_dim, output_dim, n_experts, classification, activation=None, initializer=None,
                 featLearn=0, cnn=1):
        MoE.__init__(self, name, input_type, input_dim, 1, output_dim, n_experts, classification, activation=activation,
                     initializer=initializer, lossVAE=0, featLearn=featLearn, cnn=cnn)

    def _define_vae(self):
        self.vae = DeepMixtureVAE("/".join([self.name] * 3) + "/null_vae", self.input_type, self.input_dim, self.latent_dim,
                                  self.n_experts, activation=self.activation, initializer=self.initializer).build_graph()


class DeepVariationalMoE(MoE):
    """models.py:252-262."""

    def __init__(self, name, input_type, input_dim, latent_dim, output_dim, n_experts, classification, activation=None,
                 initializer=None, featLearn=1, cnn=1):
        MoE.__init__(self, name, input_type, input_dim, latent_dim, output_dim, n_experts, classification,
                     activation=activation, initializer=initializer, featLearn=featLearn, cnn=cnn)

    def _define_vae(self):
        self.vae = DeepMixtureVAE("/".join([self.name] * 3) + "/deep_mixture_vae", self.input_type, self.input_dim,
                                  self.latent_dim, self.n_experts, activation=self.activation,
                                  initializer=self.initializer).build_graph()


class VaDEMoE(MoE):
    """models.py:265-275: the gate is gamma = get_cluster_probs(Z); training sends the supervised gradient through it."""

    def __init__(self, name, input_type, input_dim, latent_dim, output_dim, n_experts, classification, activation=None,
                 initializer=None, featLearn=1, cnn=1):
        MoE.__init__(self, name, input_type, input_dim, latent_dim, output_dim, n_experts, classification,
                     activation=activation, initializer=initializer, featLearn=featLearn, cnn=cnn)

    def _define_vae(self):
        self.vae = VaDE("/".join([self.name] * 3) + "/vade", self.input_type, self.input_dim, self.latent_dim, self.n_experts,
                        activation=self.activation, initializer=self.initializer).build_graph()
