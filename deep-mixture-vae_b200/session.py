"""Session and graph handles: the seam of the reference is ``session.run(fetches, feed_dict)``
(code/base_models.py:126).  Here a Session is a thin context object (device + stream); handles are named
references to device buffers of the model's engine."""
from __future__ import annotations

from typing import Any, Dict, Iterable, List, Optional

import numpy as np
import torch


class Handle:
    """A named graph element (placeholder, tensor or op) owned by a model."""
    __slots__ = ("name", "owner", "kind")

    def __init__(self, name: str, owner, kind: str):
        self.name, self.owner, self.kind = name, owner, kind

    def __repr__(self):
        return "<%s %s of %s>" % (self.kind, self.name, getattr(self.owner, "name", "?"))

    def __hash__(self):
        return id(self)

    def __eq__(self, other):
        return self is other


class Session:
    """Replaces tf.Session: picks the CUDA device; models bind their engines to it on first use."""

    def __init__(self, device: Optional[int] = None, gemm_dtype: Optional[str] = None, seed: int = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("dmvae_b200.Session needs a CUDA device: there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.gemm_dtype = gemm_dtype
        self.seed = seed
        self.closed = False

    def run(self, fetches, feed_dict: Optional[Dict[Handle, Any]] = None):
        single = isinstance(fetches, Handle)
        fl: List[Handle] = [fetches] if single else list(fetches)
        feed_dict = feed_dict or {}
        owners = [f.owner for f in fl] + [k.owner for k in feed_dict.keys()]
        root = None
        for o in owners:                       # an MoE wrapper outranks the VAE it owns
            if getattr(o, "is_moe", False):
                root = o
                break
        if root is None:
            root = owners[0]
        values = root._run(self, fl, feed_dict)
        return values[0] if single else values

    def close(self):
        self.closed = True

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class _Initializer:
    def run(self, session=None):
        return None


def global_variables_initializer():
    """Variables are initialised when the model builds its engine; kept so reference-style scripts run."""
    return _Initializer()
