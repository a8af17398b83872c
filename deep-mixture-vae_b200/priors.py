"""Latent variables / priors with the API of code/priors.py.

Inside a model the KL terms, the reparameterisation and q(c|x) are fused into the engine's kernels; the classes
here carry the shapes, the host-side noise samplers (same NumPy calls as the reference) and eager versions of the
maths on CUDA tensors (each a single kernel call through ``functional``), with the reference's assertion
behaviour for missing parameter keys."""
from __future__ import annotations

import numpy as np
import torch

from . import _abi, functional as F
from .includes.utils import sample_gumbel


class LatentVariable:
    def kl_from_prior(self, **kwargs):
        raise NotImplementedError

    def sample_reparametrization_variable(self, **kwargs):
        raise NotImplementedError

    def sample_generative_feed(self, **kwargs):
        raise NotImplementedError

    def inverse_reparametrize(self, **kwargs):
        raise NotImplementedError


class NormalFactorial(LatentVariable):
    """priors.py:21-47 (standard normal prior)."""

    def __init__(self, name, dim):
        self.name = name
        self.dim = dim

    def sample_reparametrization_variable(self, n):
        return np.random.randn(n, self.dim)

    def sample_generative_feed(self, n, **kwargs):
        return np.random.randn(n, self.dim)

    def inverse_reparametrize(self, epsilon, parameters):
        assert("mean" in parameters and "log_var" in parameters)
        Z, _, _ = F.reparametrize(parameters["mean"], parameters["log_var"], epsilon)
        return Z

    def kl_from_prior(self, parameters, eps=1e-20):
        """0.5 * mean_b sum_l (e^{lv} + mu^2 - 1 - lv): the mixture kernel with K = 1, m = 0, plv = 0."""
        assert("mean" in parameters and "log_var" in parameters)
        mean = parameters["mean"]
        z = torch.zeros(1, self.dim, device=mean.device)
        out = F.elbo_terms(_abi.MODE_DMVAE, mean, parameters["log_var"], z, z,
                           logits=torch.zeros(mean.shape[0], 1, device=mean.device))
        return out["per_sample"][:, 2].mean()


class NormalMixtureFactorial(LatentVariable):
    """priors.py:50-147.  ``means`` / ``log_vars`` are [K, L]; inside a model they are views of the engine's flat
    parameter buffer (trained by the fused ELBO backward), stand-alone they are local CUDA tensors."""

    def __init__(self, name, dim, n_classes, trainable=True):
        self.name = name
        self.dim = dim
        self.n_classes = n_classes
        self.trainable = trainable
        self._engine = None
        self._means = None
        self._log_vars = None

    def attach(self, engine):
        self._engine = engine

    def _local(self):
        if self._means is None:
            dev = torch.device("cuda", torch.cuda.current_device())
            self._means = torch.from_numpy(np.random.standard_normal((self.n_classes, self.dim)).astype(np.float32)).to(dev)
            self._log_vars = torch.zeros(self.n_classes, self.dim, device=dev)

    @property
    def means(self):
        if self._engine is not None:
            return self._engine.table("means")
        self._local()
        return self._means

    @property
    def log_vars(self):
        if self._engine is not None:
            return self._engine.table("log_vars")
        self._local()
        return self._log_vars

    def sample_reparametrization_variable(self, n):
        return np.random.randn(n, self.dim)

    def sample_generative_feed(self, n, **kwargs):
        """priors.py:70-84 (the session argument is accepted and unused: the tables are read directly)."""
        assert("session" in kwargs)
        samples = np.random.randn(n, self.dim)
        if "c" not in kwargs:
            c = np.random.randint(0, 10, n, dtype=np.int32)
        else:
            c = kwargs["c"]
        means = self.means.detach().cpu().numpy()[c, :]
        log_vars = self.log_vars.detach().cpu().numpy()[c, :]
        return means + samples * np.exp(log_vars / 2.0)

    def inverse_reparametrize(self, epsilon, parameters):
        assert("mean" in parameters and "log_var" in parameters)
        Z, _, _ = F.reparametrize(parameters["mean"], parameters["log_var"], epsilon)
        return Z

    def get_cluster_probs(self, Z):
        """softmax_k(-1/2 [sum_l (z-m)^2/e^{plv} + sum_l plv])  (priors.py:91-102): the VaDE mode of the fused kernel
        evaluated at mean = Z, eps = 0."""
        out = F.elbo_terms(_abi.MODE_VADE, Z, torch.zeros_like(Z), self.means, self.log_vars, eps=torch.zeros_like(Z))
        return out["qc"]

    def kl_from_prior(self, parameters, eps=1e-20):
        assert(
            "cluster_sample" in parameters and
            "weights" in parameters and
            "log_var" in parameters and
            "mean" in parameters
        )
        mean, log_var = parameters["mean"], parameters["log_var"]
        weights = parameters["weights"].reshape(-1, self.n_classes)
        if parameters["cluster_sample"]:
            out = F.elbo_terms(_abi.MODE_DMVAE_SAMPLED, mean, log_var, self.means, self.log_vars,
                               logits=torch.zeros_like(weights), zeta=weights)
        else:
            # the analytic mode weights by softmax(logits); softmax(log w) = w for a normalised w
            out = F.elbo_terms(_abi.MODE_DMVAE, mean, log_var, self.means, self.log_vars,
                               logits=torch.log(weights.to(torch.float32) + 1e-38))
        return out["per_sample"][:, 2].mean()


class DiscreteFactorial(LatentVariable):
    """priors.py:150-201."""

    def __init__(self, name, dim, n_classes):
        self.name = name
        self.dim = dim
        self.n_classes = n_classes

    def sample_reparametrization_variable(self, n):
        return sample_gumbel((n, self.dim, self.n_classes))

    def sample_generative_feed(self, n, **kwargs):
        samples = sample_gumbel((n, self.dim, self.n_classes))
        samples = np.reshape(samples, (-1, self.n_classes))
        samples = np.asarray(np.equal(samples, np.max(samples, 1, keepdims=True)), dtype=samples.dtype)
        return np.reshape(samples, (-1, self.dim, self.n_classes))

    def inverse_reparametrize(self, epsilon, parameters):
        assert("logits" in parameters and "temperature" in parameters)
        logits = parameters["logits"].reshape(-1, self.n_classes)
        B = logits.shape[0]
        z = torch.zeros(B, 4, device=logits.device)
        _, _, zeta = F.reparametrize(z, z, z, logits=logits, gumbel=epsilon, tau=float(parameters["temperature"]))
        return zeta.reshape(-1, self.dim, self.n_classes)

    def kl_from_prior(self, parameters, eps=1e-20):
        if "logits" in parameters:
            logits = parameters["logits"].reshape(-1, self.n_classes)
        elif "probs" in parameters:
            logits = torch.log(parameters["probs"].reshape(-1, self.n_classes).to(torch.float32) + 1e-38)
        else:
            assert(False)
        B = logits.shape[0]
        z = torch.zeros(B, 1, device=logits.device)
        pm = torch.zeros(self.n_classes, 1, device=logits.device)
        out = F.elbo_terms(_abi.MODE_DMVAE, z, z, pm, pm, logits=logits)
        return out["per_sample"][:, 1].mean()
