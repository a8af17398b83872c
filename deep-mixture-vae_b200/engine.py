"""Step engine: owns the flat parameter / gradient / Adam buffers and the padded activation buffers of one
model, and sequences the sm_100a kernels of libdmvae_b200 for the forward, ELBO and backward passes.

PyTorch is used for device memory and streams only; every arithmetic op on the path is a kernel of the
C-ABI library (no torch math, no CPU fallback).

Layout ("ones column", see include/dmvae_b200.h): a dense layer with n inputs and m outputs is one fp32
matrix [pad(n), pad(m)] in the flat buffer, rows [0,n) = kernel, row n = bias, everything else zero;
activations are [rows, pad(n)] with column n == 1.  pad(d) = round_up(d + 1, 64).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _abi
from ._abi import BF16, F32, U8


def pad_dim(d: int) -> int:
    return ((d + 1 + 63) // 64) * 64


def round64(d: int) -> int:
    return ((d + 63) // 64) * 64


@dataclass
class VarView:
    """Where one reference variable lives inside a padded layer matrix."""
    layer: str
    kind: str                 # "kernel" | "bias" | "table"
    col0: int = 0
    ncols: int = 0
    shape: Tuple[int, ...] = ()
    init: str = "xavier"      # xavier | zeros | normal
    trainable: bool = True
    transform: Optional[str] = None    # "moe_w" / "moe_b": reference layout differs from storage


@dataclass
class Layer:
    name: str
    n_in: int
    in_pad: int
    out_pad: int
    n_valid: int              # ones-column structure of the OUTPUT (n_valid >= n_block: none)
    n_block: int
    offset: int = 0           # float offset in the flat buffer

    @property
    def size(self) -> int:
        return self.in_pad * self.out_pad


class Layout:
    """Pure-Python description of the padded parameter layout of one model (no CUDA needed): the dense layers, where
    each reference variable lives inside them, and the offsets in the flat fp32 buffer."""

    def __init__(self, *, model, input_dim, latent_dim, n_classes, trunk, head, decoder, name, moe=None):
        if model not in ("dmvae", "vade"):
            raise NotImplementedError(model)
        self.model, self.name = model, name
        self.D, self.L, self.K = input_dim, latent_dim, n_classes
        self.Kc = ((n_classes + 7) // 8) * 8
        self.trunk, self.head, self.decoder, self.moe = tuple(trunk), head, tuple(decoder), moe
        self._build_layers()
        off = 0
        for ly in self.layers.values():
            ly.offset = off
            off += ly.size
        self.tab_size = ((self.K * self.L + 63) // 64) * 64
        self.off_means = off
        self.off_log_vars = off + self.tab_size
        self.n_params = off + 2 * self.tab_size

    def block_range(self, name: str) -> Tuple[int, int]:
        """(offset, n) of a dense layer's block, or of the two prior tables ("priors")."""
        if name == "priors":
            return self.off_means, 2 * self.tab_size
        ly = self.layers[name]
        return ly.offset, ly.size

    def stream_plan(self) -> Dict[str, List[str]]:
        """Which blocks become final together during the backward pass (key -> block names), in the order they do."""
        chain = self.dec_chain
        plan: Dict[str, List[str]] = {}
        top = ["decx", "priors"]
        if len(chain) > 1:
            plan["dec"] = list(chain[1:]) + top          # the decoder above its first layer (+ the prior tables)
            top = []
        plan["heads"] = (["zh", "ch"] if self.model == "dmvae" else ["zh"]) + [chain[0]] + top
        if self.model == "dmvae":
            plan["ench"] = ["ench"]
        for nm in self.enc_chain[1:]:
            plan["enc:" + nm] = [nm]
        return plan

    def merged_ranges(self, names) -> List[Tuple[int, int]]:
        """(offset, n) of the blocks `names`, adjacent blocks merged."""
        merged: List[Tuple[int, int]] = []
        for off, n in sorted(self.block_range(nm) for nm in names):
            if merged and merged[-1][0] + merged[-1][1] == off:
                merged[-1] = (merged[-1][0], merged[-1][1] + n)
            else:
                merged.append((off, n))
        return merged

    def stream_partition(self) -> List[Tuple[int, int]]:
        """[lo, hi) ranges tiling the flat buffer: one per streamed segment plus whatever lies between them (the first
        encoder layer, expert blocks).  The data-parallel exchange shards each range separately."""
        rs = sorted(r for names in self.stream_plan().values() for r in self.merged_ranges(names))
        out, pos = [], 0
        for lo, n in rs:
            if lo > pos:
                out.append((pos, lo))
            out.append((lo, lo + n))
            pos = lo + n
        if pos < self.n_params:
            out.append((pos, self.n_params))
        return out

    def replicated_master_ranges(self) -> List[Tuple[int, int]]:
        """[lo, hi) ranges of the flat buffer whose fp32 master every data-parallel rank needs locally: the prior tables
        (read by the fused ELBO kernel) and the logits layer (its split bf16 operand copy is rebuilt from the master)."""
        rs = []
        if "ch" in self.layers:
            ly = self.layers["ch"]
            rs.append((ly.offset, ly.offset + ly.size))
        rs.append((self.off_means, self.n_params))
        return rs

    def reference_parameter_count(self) -> int:
        """Number of reference parameters that receive a gradient (4 373 014 for cfg1/2)."""
        return int(sum(int(np.prod(v.shape)) for v in self.vars.values()))

    def _build_layers(self):
        D, L, K, n = self.D, self.L, self.K, self.name
        self.layers: Dict[str, Layer] = {}
        self.vars: Dict[str, VarView] = {}
        e, d = n + "/encoder_network", n + "/decoder_network"

        def add(name, n_in, n_out_pad, n_valid, n_block):
            self.layers[name] = Layer(name, n_in, pad_dim(n_in), n_out_pad, n_valid, n_block)

        def dense_vars(layer, kname, bname, n_in, col0, ncols, bias_shape, bias_init):
            self.vars[kname] = VarView(layer, "kernel", col0, ncols, (n_in, ncols), "xavier")
            self.vars[bname] = VarView(layer, "bias", col0, ncols, bias_shape, bias_init)

        if self.model == "dmvae":
            h1, h2 = self.trunk
            hh = self.head
            add("enc1", D, pad_dim(h1), h1, pad_dim(h1))
            dense_vars("enc1", e + "/dense/kernel", e + "/dense/bias", D, 0, h1, (h1,), "zeros")
            add("enc2", h1, pad_dim(h2), h2, pad_dim(h2))
            dense_vars("enc2", e + "/dense_1/kernel", e + "/dense_1/bias", h1, 0, h2, (h2,), "zeros")
            hp = pad_dim(hh)
            add("ench", h2, 2 * hp, hh, hp)                               # [hidden_z | hidden_c]
            dense_vars("ench", e + "/z/dense/kernel", e + "/z/dense/bias", h2, 0, hh, (hh,), "zeros")
            dense_vars("ench", e + "/c/dense/kernel", e + "/c/dense/bias", h2, hp, hh, (hh,), "zeros")
            add("zh", hh, round64(2 * L), 1, 1)                           # [mean | log_var], fp32 out
            dense_vars("zh", e + "/z/dense_1/kernel", e + "/z/dense_1/bias", hh, 0, L, (L,), "zeros")
            dense_vars("zh", e + "/z/dense_2/kernel", e + "/z/dense_2/bias", hh, L, L, (L,), "zeros")
            # logits, fp32 out.  Three column groups of stride Kc: the bf16 operand copy carries W = hi + lo + lo2 there
            # (dmvae_split3_bf16) so that the bf16 tier's logits are fp32-exact; the fp32 master has K valid columns
            self.Kc = ((K + 7) // 8) * 8
            add("ch", hh, round64(3 * self.Kc), 1, 1)
            dense_vars("ch", e + "/c/dense_1/kernel", e + "/c/dense_1/bias", hh, 0, K, (K,), "zeros")
            self.enc_chain = ["enc1", "enc2"]
            self.last_hidden = h2
            # dead head reconstructed_Y_soft (base_models.py:251-253): kept as untrained host variables
            self.dead_vars = {e + "/dense_2/kernel": (h2, 10), e + "/dense_2/bias": (10,)}
        else:
            prev = D
            self.enc_chain = []
            for i, h in enumerate(self.trunk):
                nm = "enc%d" % (i + 1)
                add(nm, prev, pad_dim(h), h, pad_dim(h))
                dense_vars(nm, e + "/layers/layer_%d/weight" % (i + 1), e + "/layers/layer_%d/bias" % (i + 1), prev, 0,
                           h, (1, h), "xavier")                          # FullyConnected bias is xavier (layers.py:27-28)
                self.enc_chain.append(nm)
                prev = h
            add("zh", prev, round64(2 * L), 1, 1)
            dense_vars("zh", e + "/z/dense/kernel", e + "/z/dense/bias", prev, 0, L, (L,), "zeros")
            dense_vars("zh", e + "/z/dense_1/kernel", e + "/z/dense_1/bias", prev, L, L, (L,), "zeros")
            self.last_hidden = prev
            self.dead_vars = {}
        prev = L
        self.dec_chain = []
        for i, h in enumerate(self.decoder):
            nm = "dec%d" % (i + 1)
            add(nm, prev, pad_dim(h), h, pad_dim(h))
            dense_vars(nm, d + "/layers/layer_%d/weight" % (i + 1), d + "/layers/layer_%d/bias" % (i + 1), prev, 0, h,
                       (1, h), "xavier")
            self.dec_chain.append(nm)
            prev = h
        add("decx", prev, pad_dim(D), 1, 1)
        dense_vars("decx", d + "/dense/kernel", d + "/dense/bias", prev, 0, D, (D,), "zeros")
        if self.moe is not None:
            E, O = self.moe["n_experts"], self.moe["output_dim"]
            I = L if self.moe["featLearn"] else D
            add("moe", I, round64(E * O), 1, 1)
            pfx = self.moe["scope"]
            self.vars[pfx + "/regression_weights"] = VarView("moe", "kernel", 0, E * O, (E, O, I), "normal", True, "moe_w")
            self.vars[pfx + "/regression_biases"] = VarView("moe", "bias", 0, E * O, (O, E), "zeros", True, "moe_b")
        # prior tables (priors.py:58-65) live at the end of the flat buffer
        self.vars[n + "/representation/means"] = VarView("prior_means", "table", 0, L, (K, L), "normal")
        self.vars[n + "/representation/log_vars"] = VarView("prior_log_vars", "table", 0, L, (K, L), "zeros")


_TORCH_DT = {F32: torch.float32, BF16: torch.bfloat16, U8: torch.uint8}


class AdamState:
    """Slots of one tf.train.AdamOptimizer instance (base_models.py:102-110, :307-321)."""

    def __init__(self, n: int, device, lr: float, beta1=0.9, beta2=0.999, eps=1e-8):
        self.m = torch.zeros(n, dtype=torch.float32, device=device)
        self.v = torch.zeros(n, dtype=torch.float32, device=device)
        self.t = 0
        self.lr, self.beta1, self.beta2, self.eps = lr, beta1, beta2, eps
        # device-resident {uint64 step; uint32 t; float lr_t} read by the kernels under CUDA-graph replay
        self.state_dev = torch.zeros(4, dtype=torch.int32, device=device)
        self.state_valid = False

    def next_lr_t(self) -> float:
        self.t += 1
        self.state_valid = False
        return self.lr * math.sqrt(1.0 - self.beta2 ** self.t) / (1.0 - self.beta1 ** self.t)

    def upload_state(self, step_count: int):
        """device step = step_count - 1 and t = self.t: the graph's first node (dmvae_step_tick) increments both."""
        st = np.zeros(1, dtype=[("step", "<u8"), ("t", "<u4"), ("lr_t", "<f4")])
        st["step"] = (step_count - 1) & 0xFFFFFFFFFFFFFFFF
        st["t"] = self.t
        self.state_dev.copy_(torch.from_numpy(st.view(np.int32).copy()), non_blocking=False)
        self.state_valid = True


class Engine:
    """Kernels + buffers for one DMVAE / VaDE model (optionally with an MoE expert head)."""

    def __init__(self, *, model: str, input_type: str, input_dim: int, latent_dim: int, n_classes: int,
                 trunk: Tuple[int, ...], head: int, decoder: Tuple[int, ...], name: str,
                 gemm_dtype: str = "bf16", device=None, seed: int = 0, max_rows: int = 4096,
                 cluster_sample: bool = False, temperature: float = 1.0, decoded_dtype: Optional[str] = None,
                 moe: Optional[dict] = None, split_k_wgrad: Optional[int] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("dmvae_b200 needs a CUDA device: there is no CPU fallback")
        if input_type not in ("binary", "real"):
            raise NotImplementedError(input_type)                      # base_models.py:84-85
        if model not in ("dmvae", "vade"):
            raise NotImplementedError(model)
        self.lib = _abi.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.model, self.input_type, self.name = model, input_type, name
        self.D, self.L, self.K = input_dim, latent_dim, n_classes
        self.trunk, self.head, self.decoder = tuple(trunk), head, tuple(decoder)
        self.cluster_sample, self.temperature = cluster_sample, float(temperature)
        self.dt = {"bf16": BF16, "fp32": F32, "f32": F32}[gemm_dtype]
        self.tdt = _TORCH_DT[self.dt]
        self.dec_dt = self.dt if decoded_dtype is None else {"bf16": BF16, "fp32": F32, "f32": F32}[decoded_dtype]
        if self.dt == F32:
            self.dec_dt = F32
        self.moe = moe
        # exact cluster assignments in the bf16 tier: split-weight logits layer (include/dmvae_b200.h, dmvae_split3_bf16)
        self.split_heads = (self.dt == BF16 and model == "dmvae" and os.environ.get("DMVAE_SPLIT_HEADS", "1") != "0")
        self.seed = seed
        self.noise_seed = seed + 2
        self.step_count = 0
        self.split_k_wgrad = split_k_wgrad
        self.world, self.rank = 1, 0
        self.dp = None
        self._graphs = {}
        self._graph_replay_launches = 0
        self.klr_dev = torch.ones(1, dtype=torch.float32, device=self.device)
        self._klr_host = 1.0
        ctx = C.c_void_p()
        _abi.check(self.lib.dmvae_ctx_create(self.device.index or 0, C.byref(ctx)))
        self.ctx = ctx
        if self.dt == BF16 and not self.lib.dmvae_ctx_has_tcgen05(self.ctx):
            raise RuntimeError("gemm_dtype='bf16' needs an sm_100 (B200) device with TMA; no fallback exists")
        self.layout = Layout(model=model, input_dim=input_dim, latent_dim=latent_dim, n_classes=n_classes, trunk=trunk,
                             head=head, decoder=decoder, name=name, moe=moe)
        for k in ("layers", "vars", "enc_chain", "dec_chain", "dead_vars", "last_hidden", "tab_size", "off_means",
                  "off_log_vars", "n_params", "Kc"):
            setattr(self, k, getattr(self.layout, k))
        self._alloc_params()
        self.max_rows = 0
        self._alloc_activations(max_rows)
        self.init_variables(seed)

    # ------------------------------------------------------------------------------------------
    # layer table
    # ------------------------------------------------------------------------------------------
    def _alloc_params(self):
        dev = self.device
        self.params = torch.zeros(self.n_params, dtype=torch.float32, device=dev)
        self.grads = torch.zeros(self.n_params, dtype=torch.float32, device=dev)
        self.params_op = torch.zeros(self.n_params, dtype=torch.bfloat16, device=dev) if self.dt == BF16 else None
        self.optimizers: Dict[str, AdamState] = {}
        self.dead_values = {}

    def W(self, name: str, grad: bool = False, op: bool = False) -> torch.Tensor:
        ly = self.layers[name]
        buf = self.grads if grad else (self.params_op if (op and self.params_op is not None) else self.params)
        return buf[ly.offset: ly.offset + ly.size].view(ly.in_pad, ly.out_pad)

    def table(self, which: str, grad: bool = False) -> torch.Tensor:
        off = self.off_means if which == "means" else self.off_log_vars
        buf = self.grads if grad else self.params
        return buf[off: off + self.K * self.L].view(self.K, self.L)

    # ------------------------------------------------------------------------------------------
    # variables (reference names)
    # ------------------------------------------------------------------------------------------
    def variable_names(self) -> List[str]:
        return list(self.vars.keys()) + list(self.dead_vars.keys())

    def trainable_size(self) -> int:
        return self.layout.reference_parameter_count()

    def _view(self, vv: VarView, grad=False) -> torch.Tensor:
        if vv.kind == "table":
            return self.table("means" if vv.layer == "prior_means" else "log_vars", grad)
        ly = self.layers[vv.layer]
        Wm = self.W(vv.layer, grad)
        if vv.kind == "kernel":
            return Wm[: ly.n_in, vv.col0: vv.col0 + vv.ncols]
        return Wm[ly.n_in, vv.col0: vv.col0 + vv.ncols]

    def get_variable(self, name: str, grad: bool = False) -> np.ndarray:
        if name in self.dead_vars:
            return self.dead_values[name].copy()
        if self.dp is not None and not grad:
            self.dp.gather_master()            # data parallel, bf16 tier: fp32 master shards live on their owners
        vv = self.vars[name]
        t = self._view(vv, grad).detach().cpu().numpy()
        if vv.transform == "moe_w":          # storage [I, E*O] -> reference [E, O, I]
            E, O, I = vv.shape
            return np.ascontiguousarray(t.reshape(I, E, O).transpose(1, 2, 0))
        if vv.transform == "moe_b":          # storage [E*O] -> reference [O, E]
            O, E = vv.shape
            return np.ascontiguousarray(t.reshape(E, O).T)
        return t.reshape(vv.shape).copy()

    def set_variable(self, name: str, value) -> None:
        value = np.asarray(value, dtype=np.float32)
        if name in self.dead_vars:
            self.dead_values[name] = value.reshape(self.dead_vars[name]).copy()
            return
        if self.dp is not None:
            self.dp.gather_master()            # the recast below reads the whole fp32 buffer: peers' shards must be current
        vv = self.vars[name]
        if tuple(value.shape) != tuple(vv.shape):
            value = value.reshape(vv.shape)
        if vv.transform == "moe_w":
            E, O, I = vv.shape
            value = value.transpose(2, 0, 1).reshape(I, E * O)
        elif vv.transform == "moe_b":
            value = value.T.reshape(-1)
        view = self._view(vv)
        view.copy_(torch.from_numpy(np.ascontiguousarray(value)).to(self.device).reshape(view.shape))
        self._params_dirty = True

    def load_variables(self, values: Dict[str, np.ndarray], strict: bool = False) -> None:
        for k, v in values.items():
            if k in self.vars or k in self.dead_vars:
                self.set_variable(k, v)
            elif strict:
                raise KeyError(k)
        self.sync_operand_copy()

    def state_dict(self) -> Dict[str, np.ndarray]:
        return {k: self.get_variable(k) for k in self.variable_names()}

    def init_variables(self, seed: int) -> None:
        """xavier-uniform kernels, zero dense biases, xavier FullyConnected biases, N(0,1) prior means, zero prior
        log-variances (train.py:161, layers.py:25-28, priors.py:58-65).  Host RNG, one upload."""
        rng = np.random.RandomState(seed)
        for k, shape in self.dead_vars.items():
            if len(shape) == 2:
                a = math.sqrt(6.0 / (shape[0] + shape[1]))
                self.dead_values[k] = rng.uniform(-a, a, size=shape).astype(np.float32)
            else:
                self.dead_values[k] = np.zeros(shape, np.float32)
        for k, vv in self.vars.items():
            shape = vv.shape
            if vv.init == "xavier":
                fan_in, fan_out = shape[-2], shape[-1]
                a = math.sqrt(6.0 / (fan_in + fan_out))
                val = rng.uniform(-a, a, size=shape).astype(np.float32)
            elif vv.init == "normal":
                val = rng.standard_normal(shape).astype(np.float32)
            else:
                val = np.zeros(shape, np.float32)
            self.set_variable(k, val)
        self.sync_operand_copy()

    def sync_operand_copy(self) -> None:
        """Refresh the bf16 operand copy of the parameters (after loading variables)."""
        if self.dp is not None:
            self.dp.gather_master()
        if self.params_op is not None:
            _abi.check(self.lib.dmvae_cast_bf16(self.ctx, self.params.data_ptr(), self.params_op.data_ptr(),
                                                self.n_params, self._stream()))
            self._refresh_split_heads()
        self._params_dirty = False

    def _refresh_split_heads(self) -> None:
        """Rebuild the three-group bf16 operand copy (hi | lo | lo2) of the logits layer from its fp32 master; called
        after every rewrite of that block's operand copy (load, Adam, data-parallel exchange)."""
        if not self.split_heads:
            return
        ly = self.layers["ch"]
        _abi.check(self.lib.dmvae_split3_bf16(self.ctx, self.params.data_ptr() + 4 * ly.offset,
                                              self.params_op.data_ptr() + 2 * ly.offset, ly.in_pad, ly.out_pad, self.K,
                                              self.Kc, self._stream()))

    def _fold_logits(self, rows: int) -> None:
        """logits = hi + lo + lo2 partial products of the split-weight GEMM (when reparam() is not the one folding)."""
        if self.split_heads:
            _abi.check(self.lib.dmvae_fold3(self.ctx, self.ch.data_ptr(), self.ch.stride(0), rows, self.K, self.Kc,
                                            self._stream()))

    # ------------------------------------------------------------------------------------------
    # activations
    # ------------------------------------------------------------------------------------------
    def _alloc_activations(self, rows: int):
        if rows <= self.max_rows:
            return
        self.max_rows = rows
        # captured graphs and the epoch staging hold the old buffers' addresses: drop them with the buffers
        self._graphs = {}
        self._epoch_key = None
        dev, t = self.device, self.tdt
        B = rows
        z = lambda cols, dt=t: torch.zeros(B, cols, dtype=dt, device=dev)
        f32 = torch.float32
        self.act: Dict[str, torch.Tensor] = {}
        self.dact: Dict[str, torch.Tensor] = {}
        self.act["x"] = z(pad_dim(self.D))
        for nm in self.enc_chain + (["ench"] if self.model == "dmvae" else []) + self.dec_chain:
            self.act[nm] = z(self.layers[nm].out_pad)
            self.dact[nm] = z(self.layers[nm].out_pad)
        self.zh = z(self.layers["zh"].out_pad, f32)                    # mean | log_var
        self.dzh = z(self.layers["zh"].out_pad)
        if self.model == "dmvae":
            self.ch = z(self.layers["ch"].out_pad, f32)                # logits
            self.dch = z(self.layers["ch"].out_pad)
        self.zb = z(pad_dim(self.L))                                   # Z (operand dtype, ones column)
        self.dz = z(pad_dim(self.L), f32)
        self.decoded = z(pad_dim(self.D), _TORCH_DT[self.dec_dt])
        self.ddecoded = z(pad_dim(self.D), _TORCH_DT[self.dec_dt])
        self.eps = z(self.L, f32)
        self.eps_in = z(self.L, f32)
        self.gumbel_in = z(self.K, f32)
        self.zeta = z(self.K, f32)
        self.per_sample = z(4, f32)
        self.r_part = z((pad_dim(self.D) + 31) // 32, f32)             # fused reconstruction term: one slot per column range
        self.qc = z(self.K, f32)
        self.argmax = torch.zeros(B, dtype=torch.int32, device=dev)
        self.dmean_kl = z(self.L, f32)
        self.dlogvar_kl = z(self.L, f32)
        self.dz_gamma = z(self.L, f32)
        self.w_scratch = z(self.K, f32)
        self.f_scratch = z(2 * self.L, f32)
        if getattr(self, "loss_out", None) is None:
            self.loss_out = torch.zeros(4, dtype=f32, device=dev)      # allocated once: fetches keep reading the same scalar block
            # loss terms of the last LOSS_RING steps, slot = step counter % LOSS_RING (dmvae_log_append).  The ring is PINNED
            # HOST memory: every captured step writes its 16 bytes across the bus itself (the per-step device -> host read of
            # the loss), and run_epoch only has to look at them after its final synchronisation
            self.loss_ring = torch.zeros(self.LOSS_RING, 4, dtype=f32).pin_memory()
        ws = int(self.lib.dmvae_elbo_reduce_workspace(B, self.L, self.K))
        self.red_ws = torch.zeros(max(ws, 4), dtype=f32, device=dev)
        self.x_stage: Dict[int, torch.Tensor] = {}
        if self.moe is not None:
            E, O = self.moe["n_experts"], self.moe["output_dim"]
            self.moe_in = z(self.layers["moe"].in_pad)
            self.moe_pred = z(self.layers["moe"].out_pad, f32)
            self.moe_dpred = z(self.layers["moe"].out_pad)
            self.moe_dgate = z(E, f32)
            self.moe_ps = z(2, f32)
            self.moe_ysoft = z(O, f32)
            self.moe_cls = torch.zeros(B, dtype=torch.int32, device=dev)
            self.moe_dinp = z(self.layers["moe"].in_pad, f32)
            self.y_buf = z(O, f32)
            self.moe_loss = torch.zeros(2, dtype=f32, device=dev)
            self.moe_ring = torch.zeros(self.LOSS_RING, 2, dtype=f32).pin_memory()    # pinned host memory, like loss_ring

    LOSS_RING = 4096

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ---- side stream: weight-gradient GEMMs and cross-sample reductions run beside the dgrad chain ----
    overlap = True

    def _fork(self, fn):
        """Run fn() on the side stream after everything already queued on the current stream."""
        if not self.overlap or self.timers is not None:
            return fn()
        main = torch.cuda.current_stream(self.device)
        if getattr(self, "_side_stream", None) is None:
            self._side_stream = torch.cuda.Stream(device=self.device)
        ev = torch.cuda.Event()
        ev.record(main)
        self._side_stream.wait_event(ev)
        with torch.cuda.stream(self._side_stream):
            fn()
        self._forked = True

    def _join(self):
        """Make the current stream wait for the side stream."""
        if getattr(self, "_forked", False):
            torch.cuda.current_stream(self.device).wait_stream(self._side_stream)
            self._forked = False

    # CUDA-event timers on the launching stream (bench.py): timers = {"elbo": [], "gemm": [] ...}
    timers = None
    _log_step = None
    _log_moe = None

    def _tic(self, key):
        if self.timers is None or key not in self.timers:
            return None
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(torch.cuda.current_stream(self.device))
        return ev

    def _toc(self, key, start):
        if start is None:
            return
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(torch.cuda.current_stream(self.device))
        self.timers[key].append((start, ev))

    def timer_ms(self, key):
        """Per-launch durations (ms) of the recorded event pairs; call after a synchronize."""
        return [a.elapsed_time(b) for a, b in self.timers.get(key, [])]

    # ------------------------------------------------------------------------------------------
    # kernel wrappers
    # ------------------------------------------------------------------------------------------
    def _fwd(self, name, A, lda, Y, out_dt, act, rows, W=None, ldw=None, n_out=None, n_in_pad=None, split_k=1):
        ly = self.layers[name]
        Wt = self.W(name, op=True) if W is None else W
        t0 = self._tic("gemm")
        if split_k > 1:
            # narrow output, deep reduction (latent heads): k-splits reduce into the pre-zeroed fp32 output
            e = _abi.GemmEpilogue()
            e.out_dtype, e.act, e.n_valid, e.n_block, e.pad_one = F32, _abi.ACT_NONE, 1 << 30, 1 << 30, 0.0
            e.accumulate, e.split_k = 1, split_k
            _abi.check(self.lib.dmvae_gemm(self.ctx, self.dt, 0, 0, A.data_ptr(), lda, Wt.data_ptr(),
                                           ly.out_pad if ldw is None else ldw, Y.data_ptr(), Y.stride(0), rows,
                                           ly.out_pad if n_out is None else n_out,
                                           ly.in_pad if n_in_pad is None else n_in_pad, C.byref(e), self._stream()))
            self._toc("gemm", t0)
            return
        _abi.check(self.lib.dmvae_linear_fwd(self.ctx, self.dt, A.data_ptr(), lda, Wt.data_ptr(),
                                             ly.out_pad if ldw is None else ldw, Y.data_ptr(), Y.stride(0), out_dt, rows,
                                             ly.out_pad if n_out is None else n_out,
                                             ly.in_pad if n_in_pad is None else n_in_pad, act, ly.n_valid, ly.n_block,
                                             self._stream()))
        self._toc("gemm", t0)

    group_heads = os.environ.get("DMVAE_HEAD_GROUP", "1") != "0"
    head_group_split = int(os.environ.get("DMVAE_HEAD_GROUP_SPLIT", "2"))

    def _fwd_entry(self, name, A, lda, Y, rows, split_k):
        """Grouped-launch entry of a linear layer whose k-splits reduce into the cleared fp32 output Y."""
        ly = self.layers[name]
        g = _abi.ChainGemm()                     # Y[rows, out_pad] += A[rows, in_pad] . W[in_pad, out_pad]
        g.trans_a, g.trans_b = 0, 0
        g.A, g.lda, g.B, g.ldb = A.data_ptr(), lda, self.W(name, op=True).data_ptr(), ly.out_pad
        g.C, g.ldc = Y.data_ptr(), Y.stride(0)
        g.M, g.N, g.K = rows, ly.out_pad, ly.in_pad
        g.epi.out_dtype, g.epi.act, g.epi.n_valid, g.epi.n_block, g.epi.pad_one = F32, _abi.ACT_NONE, 1 << 30, 1 << 30, 0.0
        g.epi.accumulate, g.epi.split_k = 1, split_k
        g.dep[0], g.dep[1] = -1, -1
        return g

    # ---- grouped backward: the weight-gradient and data-gradient GEMMs that become runnable together go out as ONE
    #      persistent launch (dmvae_gemm_chain without dependencies) instead of one launch each on two streams: the
    #      separate launches could not overlap anyway (each CTA pair takes a whole SM pair's shared memory), and every
    #      launch paid its own prologue and tail.  DMVAE_GROUP=0 restores the launch-per-GEMM schedule. ----
    use_groups = os.environ.get("DMVAE_GROUP", "1") != "0"
    _group = None

    def _begin_group(self):
        self._group = [] if (self.use_groups and self.dt == BF16 and self.timers is None) else None

    def _flush_group(self):
        g = self._group
        if not g:
            return
        if len(g) == 1:
            g[0][1]()
        else:
            n = len(g)
            arr = (_abi.ChainGemm * n)(*[e for e, _ in g])
            _abi.check(self.lib.dmvae_gemm_chain(self.ctx, arr, n, None, 0, 0, None, self._stream()))
        self._group = []

    def _end_group(self):
        self._flush_group()
        self._group = None

    def _wgrad(self, name, A, lda, dY, lddy, rows, n_out=None, col0=0, accumulate=None):
        if self._group is not None:
            return self._wgrad_now(name, A, lda, dY, lddy, rows, n_out, col0, accumulate)
        self._fork(lambda: self._wgrad_now(name, A, lda, dY, lddy, rows, n_out, col0, accumulate))

    def _wgrad_now(self, name, A, lda, dY, lddy, rows, n_out=None, col0=0, accumulate=None):
        ly = self.layers[name]
        dW = self.W(name, grad=True)
        n_out = ly.out_pad if n_out is None else n_out
        sk = self._pick_split_k(rows, ly.in_pad, n_out)
        acc = 1 if sk > 1 else 0
        if accumulate is not None:
            acc = accumulate

        def direct():
            t0 = self._tic("gemm")
            _abi.check(self.lib.dmvae_linear_wgrad(self.ctx, self.dt, A.data_ptr(), lda, dY.data_ptr(), lddy,
                                                   dW.data_ptr() + 4 * col0, ly.out_pad, rows, ly.in_pad, n_out, acc, sk,
                                                   self._stream()))
            self._toc("gemm", t0)

        if self._group is None:
            return direct()
        g = _abi.ChainGemm()                     # dW[in_pad, n_out] (+)= X[rows, in_pad]^T . dY[rows, n_out]
        g.trans_a, g.trans_b = 1, 0
        g.A, g.lda, g.B, g.ldb = A.data_ptr(), lda, dY.data_ptr(), lddy
        g.C, g.ldc = dW.data_ptr() + 4 * col0, ly.out_pad
        g.M, g.N, g.K = ly.in_pad, n_out, rows
        g.epi.out_dtype, g.epi.act, g.epi.n_valid, g.epi.n_block = F32, _abi.ACT_NONE, 1 << 30, 1 << 30
        g.epi.accumulate, g.epi.split_k = acc, sk
        g.dep[0], g.dep[1] = -1, -1
        self._group.append((g, direct))

    def _head_split_k(self, rows) -> int:
        """k-splits of the narrow latent-head GEMMs (and the decoder's first dgrad): only when the single-CTA tiles of one
        split leave most SMs idle and the reduction is deep enough to share."""
        if self.dt != BF16 or self.timers is not None:
            return 1
        ctas = (rows + 127) // 128
        kin = self.layers["zh"].in_pad
        if ctas * 2 > 148 or kin < 1024:
            return 1
        return max(1, min(4, 148 // ctas, kin // 512))

    def _pick_split_k(self, rows, m, n) -> int:
        if self.dt != BF16:
            return 1
        if self.split_k_wgrad is not None:
            return max(1, self.split_k_wgrad)
        nkb = (rows + 63) // 64
        sms = 148
        if m >= 256 and n >= 128 and os.environ.get("DMVAE_GEMM_PAIR") != "0":
            # CTA-pair kernel (256 x 256 tiles, persistent): one work unit per SM pair, never more units than pairs
            tiles = ((m + 255) // 256) * ((n + 255) // 256)
            return max(1, min(nkb // 2, (sms // 2) // tiles))
        tiles = ((m + 127) // 128) * ((n + 127) // 128)
        sk = max(1, min(nkb, (2 * sms + tiles - 1) // tiles))
        return sk

    def _dgrad(self, name, dY, lddy, act_in, ld_act, dX, out_dt, rows, prev_valid, prev_block, W=None, ldw=None,
               n_in_pad=None, n_out_pad=None, split_k=1):
        ly = self.layers[name]
        Wt = self.W(name, op=True) if W is None else W
        n_in = ly.in_pad if n_in_pad is None else n_in_pad
        n_out = ly.out_pad if n_out_pad is None else n_out_pad

        def entry():
            g = _abi.ChainGemm()                 # dX[rows, n_in] = (dY[rows, n_out] . W[n_in, n_out]^T) * (act_in > 0)
            g.trans_a, g.trans_b = 0, 1
            g.A, g.lda, g.B, g.ldb = dY.data_ptr(), lddy, Wt.data_ptr(), ly.out_pad if ldw is None else ldw
            g.C, g.ldc = dX.data_ptr(), dX.stride(0)
            g.M, g.N, g.K = rows, n_in, n_out
            g.epi.out_dtype, g.epi.act = out_dt, _abi.ACT_NONE
            g.epi.n_valid, g.epi.n_block = prev_valid, prev_block if prev_block > 0 else n_in
            g.epi.pad_one = 0.0
            g.epi.relu_mask, g.epi.ld_mask = (act_in.data_ptr() if act_in is not None else None), ld_act
            g.epi.accumulate, g.epi.split_k = (1 if split_k > 1 else 0), split_k
            if split_k > 1:                      # partial sums reduce into the cleared fp32 output: linear epilogue only
                g.epi.n_valid = g.epi.n_block = 1 << 30
            g.dep[0], g.dep[1] = -1, -1
            return g

        def direct():
            t0 = self._tic("gemm")
            if split_k > 1:
                g = entry()
                _abi.check(self.lib.dmvae_gemm(self.ctx, self.dt, 0, 1, g.A, g.lda, g.B, g.ldb, g.C, g.ldc, g.M, g.N, g.K,
                                               C.byref(g.epi), self._stream()))
            else:
                _abi.check(self.lib.dmvae_linear_dgrad(self.ctx, self.dt, dY.data_ptr(), lddy, Wt.data_ptr(),
                                                       ly.out_pad if ldw is None else ldw,
                                                       act_in.data_ptr() if act_in is not None else None, ld_act,
                                                       dX.data_ptr(), dX.stride(0), out_dt, rows, n_in, n_out, prev_valid,
                                                       prev_block, self._stream()))
            self._toc("gemm", t0)

        if self._group is None:
            return direct()
        self._group.append((entry(), direct))

    # ------------------------------------------------------------------------------------------
    # input staging
    # ------------------------------------------------------------------------------------------
    def stage_input(self, X: torch.Tensor, rows: int) -> Tuple[torch.Tensor, int]:
        """X: device tensor [rows, D] (float32 / uint8 / bfloat16).  Fills the operand matrix act['x']."""
        xdt = {torch.float32: F32, torch.uint8: U8, torch.bfloat16: BF16}[X.dtype]
        _abi.check(self.lib.dmvae_stage_input(self.ctx, X.data_ptr(), xdt, X.stride(0), self.act["x"].data_ptr(), self.dt,
                                              self.act["x"].stride(0), rows, self.D, self.x_scale, self._stream()))
        return X, xdt

    # value of one unit of a uint8 input: 1 for binarised data stored as 0/1, 1/255 for 8-bit intensities (Dataset sets it
    # from its storage format); float inputs are values already and ignore it
    x_scale = 1.0

    # ------------------------------------------------------------------------------------------
    # forward
    # ------------------------------------------------------------------------------------------
    def encode(self, rows: int, heads=("z", "c"), fold_in_reparam: bool = False):
        """Encoder trunk and heads (base_models.py:218-249 / :490-513).  fold_in_reparam: the caller runs reparam() next,
        which folds the split-weight logits itself (one launch fewer on the step's critical path)."""
        self._logits_unfolded = False
        self._encode(rows, heads)
        if fold_in_reparam:
            self._logits_unfolded = self.model == "dmvae" and "c" in heads and self.split_heads
            return                               # reparam() joins the side stream (the c-head may still be running there)
        self._join()                             # stand-alone use (evaluation, fetches): both heads complete on return
        if self.model == "dmvae" and "c" in heads:
            self._fold_logits(rows)

    def _encode(self, rows: int, heads):
        a = self.act["x"]
        # A narrow head (N = 64) over a deep reduction is bound by what ONE SM can pull from L2 (~65 GB/s measured:
        # 32 CTAs took 12 us); split-K spreads the same bytes over 4x the SMs.  The partial sums reduce into fp32
        # outputs that are cleared on the side stream while the trunk runs (so is dZ, for the decoder's first dgrad).
        sk = self._head_split_k(rows)
        if sk > 1:
            def clear():
                for t in (self.zh, self.dz) + ((self.ch,) if self.model == "dmvae" else ()):
                    _abi.check(self.lib.dmvae_zero_f32(self.ctx, t.data_ptr(), rows * t.stride(0), self._stream()))
            self._fork(clear)
            self._dz_cleared = True
        for nm in self.enc_chain:
            self._fwd(nm, a, a.stride(0), self.act[nm], self.dt, _abi.ACT_RELU, rows)
            a = self.act[nm]
        if self.model == "dmvae":
            hp = self.layers["ench"].n_block
            self._fwd("ench", a, a.stride(0), self.act["ench"], self.dt, _abi.ACT_RELU, rows)
            h = self.act["ench"]
            if sk > 1:
                self._join()
            if "c" in heads and "z" in heads and sk > 1 and self.group_heads and self.use_groups and self.timers is None:
                # both heads as ONE grouped launch of the CTA-pair kernel: 2 heads x 16 row blocks x k-splits work units in
                # a single wave over the 74 SM pairs (the two single-CTA launches on two streams took two waves: each of
                # their CTAs holds a whole SM's shared memory)
                gs = self.head_group_split
                arr = (_abi.ChainGemm * 2)(self._fwd_entry("zh", h, h.stride(0), self.zh, rows, gs),
                                           self._fwd_entry("ch", h[:, hp:], h.stride(0), self.ch, rows, gs))
                _abi.check(self.lib.dmvae_gemm_chain(self.ctx, arr, 2, None, 0, 0, None, self._stream()))
                return
            if "c" in heads:
                if "z" in heads:
                    self._fork(lambda: self._fwd("ch", h[:, hp:], h.stride(0), self.ch, F32, _abi.ACT_NONE, rows, split_k=sk))
                else:
                    self._fwd("ch", h[:, hp:], h.stride(0), self.ch, F32, _abi.ACT_NONE, rows, split_k=sk)
            if "z" in heads:
                self._fwd("zh", h, h.stride(0), self.zh, F32, _abi.ACT_NONE, rows, split_k=sk)
        else:
            if sk > 1:
                self._join()
            self._fwd("zh", a, a.stride(0), self.zh, F32, _abi.ACT_NONE, rows, split_k=sk)

    def reparam(self, rows: int, eps_injected: bool, gumbel_injected: bool, row_offset: int = 0, step: Optional[int] = None,
                step_dev: Optional[int] = None):
        self._join()                 # the step-tick / output-clearing work forked at the start of the step
        ra = _abi.ReparamArgs()
        ra.step_dev = step_dev
        ra.rows, ra.L, ra.K = rows, self.L, self.K
        ra.mean = self.zh.data_ptr()
        ra.log_var = self.zh.data_ptr() + 4 * self.L
        ra.ld_zh = self.zh.stride(0)
        want_zeta = self.model == "dmvae" and self.cluster_sample
        if want_zeta:
            self._join()
        ra.logits = self.ch.data_ptr() if want_zeta else None
        ra.ld_logits = self.ch.stride(0) if want_zeta else 0
        ra.eps_in = self.eps_in.data_ptr() if eps_injected else None
        ra.gumbel_in = self.gumbel_in.data_ptr() if (gumbel_injected and want_zeta) else None
        ra.seed, ra.step, ra.row_offset = self.noise_seed, self.step_count if step is None else step, row_offset
        ra.tau = self.temperature
        ra.Z_out, ra.z_dtype, ra.ld_z, ra.z_cols = self.zb.data_ptr(), self.dt, self.zb.stride(0), self.zb.shape[1]
        ra.eps_out = self.eps.data_ptr()
        ra.zeta_out = self.zeta.data_ptr() if want_zeta else None
        if getattr(self, "_logits_unfolded", False):
            self._join()
            ra.fold, ra.ld_fold, ra.fold_K, ra.fold_stride = self.ch.data_ptr(), self.ch.stride(0), self.K, self.Kc
            self._logits_unfolded = False
        _abi.check(self.lib.dmvae_reparam_fwd(self.ctx, C.byref(ra), self._stream()))

    def decode(self, rows: int, recon=None):
        """Decoder MLP.  recon = (X, x_dtype, scale): the output layer's epilogue computes the reconstruction term and
        writes its gradient to d_decoded instead of the logits to decoded (dmvae_recon_fuse)."""
        a = self.zb
        for nm in self.dec_chain:
            self._fwd(nm, a, a.stride(0), self.act[nm], self.dt, _abi.ACT_RELU, rows)
            a = self.act[nm]
        if recon is None:
            self._fwd("decx", a, a.stride(0), self.decoded, self.dec_dt, _abi.ACT_NONE, rows)
            return
        X, xdt, scale = recon
        ly = self.layers["decx"]
        rf = _abi.ReconFuse()
        rf.X, rf.x_dtype, rf.ldx, rf.x_scale = X.data_ptr(), xdt, X.stride(0), self.x_scale
        rf.input_type = _abi.INPUT_BINARY if self.input_type == "binary" else _abi.INPUT_REAL
        rf.scale, rf.D = scale, self.D
        rf.r_part, rf.r_parts = self.r_part.data_ptr(), self.r_part.shape[1]
        e = _abi.GemmEpilogue()
        e.out_dtype, e.act, e.n_valid, e.n_block, e.pad_one, e.split_k = BF16, _abi.ACT_NONE, ly.n_valid, ly.n_block, 1.0, 1
        e.recon = C.pointer(rf)
        t0 = self._tic("gemm")
        _abi.check(self.lib.dmvae_gemm(self.ctx, self.dt, 0, 0, a.data_ptr(), a.stride(0), self.W("decx", op=True).data_ptr(),
                                       ly.out_pad, self.ddecoded.data_ptr(), self.ddecoded.stride(0), rows, ly.out_pad,
                                       ly.in_pad, C.byref(e), self._stream()))
        self._toc("gemm", t0)

    # The reconstruction term in the output layer's epilogue (SURVEY 8 row "fused ELBO"): the decoder logits never reach
    # HBM, the latent part of the ELBO runs beside the decoder's forward pass.  DMVAE_FUSE_RECON=0 turns it off.
    fuse_recon = os.environ.get("DMVAE_FUSE_RECON", "1") != "0"

    def _fuse_ok(self, X: torch.Tensor, xdt: int, rows: int = 1 << 30) -> bool:
        if rows < 256 or pad_dim(self.D) < 128:        # the CTA-pair forward kernel carries the fused epilogue
            return False
        if self.dt != BF16 or self.dec_dt != BF16 or (self.model == "dmvae" and self.cluster_sample):
            return False
        if self.K > 128 or self.L > 128:             # the latent-only kernels (row tile, split-tf32 MMA) stop there
            return False
        unit = {_abi.U8: 16, F32: 4}.get(xdt)
        return (unit is not None and self.D % unit == 0 and X.stride(0) % unit == 0 and X.data_ptr() % 16 == 0
                and X.stride(1) == 1)

    def reconstruct(self, rows: int) -> torch.Tensor:
        """reconstructed_X (base_models.py:295-300) of the rows last decoded: the output layer once more with the sigmoid
        in the GEMM epilogue (binary inputs; identity for real-valued ones).  Returns the device buffer [rows, >= D]."""
        if self.input_type != "binary":
            return self.decoded
        if getattr(self, "recon", None) is None or self.recon.shape[0] < self.max_rows:
            self.recon = torch.zeros_like(self.decoded)
        a = self.act[self.dec_chain[-1]]
        self._fwd("decx", a, a.stride(0), self.recon, self.dec_dt, _abi.ACT_SIGMOID, rows)
        return self.recon

    def decode_latent(self, Z: torch.Tensor, rows: int):
        """Generation entry point (includes/visualization.py:83-87 feeds model.Z directly): Z [rows, L] fp32 on the device
        -> decoder -> decoded_X logits in self.decoded."""
        if rows > self.max_rows:
            self._alloc_activations(rows)
        if getattr(self, "_params_dirty", False):
            self.sync_operand_copy()
        _abi.check(self.lib.dmvae_stage_features(self.ctx, Z.data_ptr(), Z.stride(0), rows, self.L, 0, self.zb.data_ptr(),
                                                 self.dt, self.zb.stride(0), self.zb.shape[1], self._stream()))
        self.decode(rows)

    # ------------------------------------------------------------------------------------------
    # chained forward: encoder -> heads (+ fused reparameterisation) -> decoder in ONE persistent launch
    # ------------------------------------------------------------------------------------------
    # off by default: measured on B200 (scripts/chain_bench.py) the chained forward is bit-identical but slower than
    # one launch per layer (115 vs 95 us at 4096 rows): the per-layer hand-off through global memory (TMA-store
    # completion -> release counter -> acquire -> TMA load) costs about as much as a PDL kernel boundary and is paid
    # on every row block's 9-layer critical path.  DMVAE_CHAIN=1 turns it on.
    use_chain = os.environ.get("DMVAE_CHAIN", "0") == "1"

    def _chain_ok(self, gumbel_injected: bool) -> bool:
        return (self.use_chain and not self.split_heads and self.dt == BF16 and self.timers is None and 2 * self.L <= 32
                and not (self.model == "dmvae" and self.cluster_sample) and not gumbel_injected)

    def _chain_entry(self, name, A, lda, C, ldc, out_dt, act, rows, dep=-1, fuse=0, n=None, k=None, W=None, ldw=None):
        ly = self.layers[name]
        g = _abi.ChainGemm()
        g.trans_a, g.trans_b = 0, 0
        Wt = self.W(name, op=True) if W is None else W
        g.A, g.lda = A.data_ptr(), lda
        g.B, g.ldb = Wt.data_ptr(), ly.out_pad if ldw is None else ldw
        g.C, g.ldc = C.data_ptr(), ldc
        g.M, g.N, g.K = rows, ly.out_pad if n is None else n, ly.in_pad if k is None else k
        g.epi.out_dtype, g.epi.act = out_dt, act
        g.epi.n_valid, g.epi.n_block = ly.n_valid, ly.n_block if ly.n_block > 0 else g.N
        g.epi.pad_one, g.epi.split_k = 1.0, 1
        g.dep[0], g.dep[1] = dep, -1
        g.fuse = fuse
        return g

    def _run_chain(self, entries, rows, ra=None, zero=True):
        n = len(entries)
        need = int(self.lib.dmvae_gemm_chain_counters(n, max(rows, self.max_rows)))
        if getattr(self, "_chain_cnt", None) is None or self._chain_cnt.numel() < need:
            self._chain_cnt = torch.zeros(need, dtype=torch.int32, device=self.device)
        arr = (_abi.ChainGemm * n)(*entries)
        _abi.check(self.lib.dmvae_gemm_chain(self.ctx, arr, n, self._chain_cnt.data_ptr(), self._chain_cnt.numel(),
                                             1 if zero else 0, C.byref(ra) if ra is not None else None, self._stream()))

    def forward_chain(self, rows: int, eps_injected: bool, row_offset: int = 0, step: Optional[int] = None,
                      step_dev: Optional[int] = None):
        """encode + reparam + decode (base_models.py:218-293) as one dmvae_gemm_chain launch."""
        self._join()                 # the step-tick forked at the start of the step (the fused epilogue reads the step)
        relu, none = _abi.ACT_RELU, _abi.ACT_NONE
        ent, idx = [], {}
        a, prev = self.act["x"], -1
        for nm in self.enc_chain:
            ent.append(self._chain_entry(nm, a, a.stride(0), self.act[nm], self.act[nm].stride(0), self.dt, relu, rows, prev))
            prev = len(ent) - 1
            a = self.act[nm]
        if self.model == "dmvae":
            h = self.act["ench"]
            hp = self.layers["ench"].n_block
            ent.append(self._chain_entry("ench", a, a.stride(0), h, h.stride(0), self.dt, relu, rows, prev))
            ih = len(ent) - 1
            ent.append(self._chain_entry("zh", h, h.stride(0), self.zh, self.zh.stride(0), F32, none, rows, ih, fuse=1))
            iz = len(ent) - 1
            ent.append(self._chain_entry("ch", h[:, hp:], h.stride(0), self.ch, self.ch.stride(0), F32, none, rows, ih))
        else:
            ent.append(self._chain_entry("zh", a, a.stride(0), self.zh, self.zh.stride(0), F32, none, rows, prev, fuse=1))
            iz = len(ent) - 1
        a, prev = self.zb, iz
        for nm in self.dec_chain:
            ent.append(self._chain_entry(nm, a, a.stride(0), self.act[nm], self.act[nm].stride(0), self.dt, relu, rows, prev))
            prev = len(ent) - 1
            a = self.act[nm]
        ent.append(self._chain_entry("decx", a, a.stride(0), self.decoded, self.decoded.stride(0), self.dec_dt, none, rows, prev))
        ra = _abi.ReparamArgs()
        ra.step_dev = step_dev
        ra.rows, ra.L, ra.K = rows, self.L, self.K
        ra.mean, ra.log_var, ra.ld_zh = self.zh.data_ptr(), self.zh.data_ptr() + 4 * self.L, self.zh.stride(0)
        ra.eps_in = self.eps_in.data_ptr() if eps_injected else None
        ra.seed, ra.step, ra.row_offset = self.noise_seed, self.step_count if step is None else step, row_offset
        ra.tau = self.temperature
        ra.Z_out, ra.z_dtype, ra.ld_z, ra.z_cols = self.zb.data_ptr(), self.dt, self.zb.stride(0), self.zb.shape[1]
        ra.eps_out = self.eps.data_ptr()
        self._run_chain(ent, rows, ra)

    def _elbo_args(self, X, xdt, rows, kl_ratio, inv_global_batch, recon_scale=1.0, klr_dev=None) -> _abi.ElboArgs:
        ea = _abi.ElboArgs()
        ea.kl_ratio_dev = klr_dev
        ea.mode = _abi.MODE_VADE if self.model == "vade" else (
            _abi.MODE_DMVAE_SAMPLED if self.cluster_sample else _abi.MODE_DMVAE)
        ea.input_type = _abi.INPUT_BINARY if self.input_type == "binary" else _abi.INPUT_REAL
        ea.rows, ea.D, ea.L, ea.K = rows, self.D, self.L, self.K
        ea.X, ea.x_dtype, ea.ldx = X.data_ptr(), xdt, X.stride(0)
        ea.decoded, ea.dec_dtype, ea.ld_dec = self.decoded.data_ptr(), self.dec_dt, self.decoded.stride(0)
        ea.mean, ea.log_var, ea.ld_zh = self.zh.data_ptr(), self.zh.data_ptr() + 4 * self.L, self.zh.stride(0)
        if self.model == "dmvae":
            ea.logits, ea.ld_logits = self.ch.data_ptr(), self.ch.stride(0)
            ea.d_logits, ea.dlogits_dtype = self.dch.data_ptr(), self.dt
            ea.ld_dlogits, ea.dlogits_cols = self.dch.stride(0), self.dch.shape[1]
        ea.eps, ea.ld_eps = self.eps.data_ptr(), self.eps.stride(0)
        ea.zeta, ea.ld_zeta = self.zeta.data_ptr(), self.zeta.stride(0)
        ea.tau = self.temperature
        ea.prior_means = self.table("means").data_ptr()
        ea.prior_log_vars = self.table("log_vars").data_ptr()
        ea.kl_ratio, ea.inv_global_batch, ea.recon_scale = kl_ratio, inv_global_batch, recon_scale
        ea.per_sample, ea.qc, ea.argmax = self.per_sample.data_ptr(), self.qc.data_ptr(), self.argmax.data_ptr()
        ea.d_decoded, ea.ld_ddec, ea.ddec_cols = self.ddecoded.data_ptr(), self.ddecoded.stride(0), self.ddecoded.shape[1]
        ea.d_mean_kl, ea.d_log_var_kl, ea.ld_dkl = self.dmean_kl.data_ptr(), self.dlogvar_kl.data_ptr(), self.L
        ea.d_Z_gamma, ea.ld_dzg = self.dz_gamma.data_ptr(), self.L
        ea.w_scratch, ea.f_scratch = self.w_scratch.data_ptr(), self.f_scratch.data_ptr()
        ea.x_scale = self.x_scale
        return ea

    def elbo_latent_fork(self, X, xdt, rows, kl_ratio=1.0, inv_global_batch=None, recon_scale=1.0, klr_dev=None,
                         prior_grads=True):
        """Fused-reconstruction step, first half: clear the r_part slots and run the LATENT part of the ELBO and the
        prior-table partial sums on the side stream (they need the encoder heads and the noise only, so they overlap
        the decoder).  Returns the argument block elbo(fused=...) completes the pass with."""
        s = (1.0 / rows) if inv_global_batch is None else inv_global_batch
        ea = self._elbo_args(X, xdt, rows, kl_ratio, s, recon_scale, klr_dev)
        ea.r_part, ea.r_parts = self.r_part.data_ptr(), self.r_part.shape[1]
        gm = self.table("means", grad=True).data_ptr() if prior_grads else None
        gl = self.table("log_vars", grad=True).data_ptr() if prior_grads else None

        def run():
            _abi.check(self.lib.dmvae_zero_f32(self.ctx, self.r_part.data_ptr(), rows * self.r_part.shape[1], self._stream()))
            t0 = self._tic("elbo")
            _abi.check(self.lib.dmvae_elbo_fwd_bwd(self.ctx, C.byref(ea), self._stream()))
            self._toc("elbo", t0)
            _abi.check(self.lib.dmvae_elbo_reduce_stage(self.ctx, C.byref(ea), gm, gl, 0, self.loss_out.data_ptr(),
                                                        self.red_ws.data_ptr(), 1, self._stream()))
        self._fork(run)
        return ea

    def elbo(self, X, xdt, rows, kl_ratio=1.0, inv_global_batch=None, recon_scale=1.0, prior_grads=True, klr_dev=None,
             d_gate_extra=None, fused=None):
        """Fused ELBO forward + backward and its cross-sample reductions."""
        s = (1.0 / rows) if inv_global_batch is None else inv_global_batch
        ea = fused if fused is not None else self._elbo_args(X, xdt, rows, kl_ratio, s, recon_scale, klr_dev)
        if d_gate_extra is not None:
            ea.d_gate_extra, ea.ld_dge = d_gate_extra.data_ptr(), d_gate_extra.stride(0)
        self._join()
        if fused is None:
            t0 = self._tic("elbo")
            _abi.check(self.lib.dmvae_elbo_fwd_bwd(self.ctx, C.byref(ea), self._stream()))
            self._toc("elbo", t0)
        gm = self.table("means", grad=True).data_ptr() if prior_grads else None
        gl = self.table("log_vars", grad=True).data_ptr() if prior_grads else None
        stage = 0 if fused is None else 2               # fused: the table partials ran with the latent part
        log = self._log_step                            # (state_dev pointer | None, host step): also file the loss terms in the ring

        def reduce():
            _abi.check(self.lib.dmvae_elbo_reduce_stage(self.ctx, C.byref(ea), gm, gl, 0, self.loss_out.data_ptr(),
                                                        self.red_ws.data_ptr(), stage, self._stream()))
            if log is not None:
                _abi.check(self.lib.dmvae_log_append(self.ctx, self.loss_out.data_ptr(), 4, self.loss_ring.data_ptr(),
                                                     self.LOSS_RING, log[0], log[1], self._stream()))
        self._fork(reduce)

    # ------------------------------------------------------------------------------------------
    # backward
    # ------------------------------------------------------------------------------------------
    def backward(self, rows: int, dmean_extra: Optional[torch.Tensor] = None, train_decoder=True, train_z=True,
                 train_c=True, train_trunk=True, through_decoder=True):
        """Gradient GEMMs from d_decoded / d_logits / KL-side gradients back to every parameter."""
        dt = self.dt
        self._begin_group()
        # ---- decoder ----
        chain = self.dec_chain
        a_last = self.act[chain[-1]]
        if through_decoder and (train_decoder or train_z):
            if train_decoder:
                self._wgrad("decx", a_last, a_last.stride(0), self.ddecoded, self.ddecoded.stride(0), rows)
            ly_prev = self.layers[chain[-1]]
            self._dgrad("decx", self.ddecoded, self.ddecoded.stride(0), a_last, a_last.stride(0), self.dact[chain[-1]], dt,
                        rows, ly_prev.n_valid, ly_prev.n_block)
            self._flush_group()
            for i in range(len(chain) - 1, -1, -1):
                nm = chain[i]
                a_in = self.act[chain[i - 1]] if i > 0 else self.zb
                dy = self.dact[nm]
                if train_decoder:
                    self._wgrad(nm, a_in, a_in.stride(0), dy, dy.stride(0), rows)
                if i > 0:
                    lp = self.layers[chain[i - 1]]
                    self._dgrad(nm, dy, dy.stride(0), a_in, a_in.stride(0), self.dact[chain[i - 1]], dt, rows, lp.n_valid,
                                lp.n_block)
                else:
                    # dZ: fp32, no ReLU mask (Z is not an activation output).  With k-splits the partial sums reduce into
                    # the cleared buffer and only columns [0, L) are meaningful (they are the only ones read)
                    sk = self._head_split_k(rows) if self.layers[nm].out_pad >= 1024 else 1
                    if not (sk > 1 and getattr(self, "_dz_cleared", False)):
                        sk = 1
                    self._dz_cleared = False
                    self._dgrad(nm, dy, dy.stride(0), None, 0, self.dz, F32, rows, self.L, self.dz.shape[1], split_k=sk)
                self._flush_group()
                if i == 1:                           # the decoder above its first layer is final (+ the prior tables)
                    self._adam_segment("dec")
        # data parallel: the decoder's gradients are complete - exchange them while the encoder's backward runs
        dpo = getattr(self, "_dp_opt", None)
        if (self.dp is not None and dpo is not None and getattr(self.dp, "overlap_decoder", False) and through_decoder
                and train_decoder and self.timers is None):
            self._fork(lambda: self.dp.early(dpo[0], dpo[1]))
        # ---- reparameterisation backward (priors.py:86-89) ----
        if train_z:
            _abi.check(self.lib.dmvae_reparam_bwd(
                self.ctx, rows, self.L, self.dmean_kl.data_ptr(), self.dlogvar_kl.data_ptr(), self.L,
                self.dz.data_ptr(), self.dz.stride(0),
                self.dz_gamma.data_ptr() if self.model == "vade" else None, self.L,
                self.eps.data_ptr(), self.zh.data_ptr() + 4 * self.L, self.zh.stride(0),
                dmean_extra.data_ptr() if dmean_extra is not None else None,
                dmean_extra.stride(0) if dmean_extra is not None else 0,
                self.dzh.data_ptr(), dt, self.dzh.stride(0), self.dzh.shape[1], self._stream()))
        # ---- encoder heads ----
        if self.model == "dmvae":
            h = self.act["ench"]
            dh = self.dact["ench"]
            hp = self.layers["ench"].n_block
            hv = self.layers["ench"].n_valid
            if train_z:
                self._wgrad("zh", h, h.stride(0), self.dzh, self.dzh.stride(0), rows)
                self._dgrad("zh", self.dzh, self.dzh.stride(0), h, h.stride(0), dh, dt, rows, hv, hp)
            if train_c:
                self._wgrad("ch", h[:, hp:], h.stride(0), self.dch, self.dch.stride(0), rows)
                self._dgrad("ch", self.dch, self.dch.stride(0), h[:, hp:], h.stride(0), dh[:, hp:], dt, rows, hv, hp)
            self._flush_group()                      # both heads' four GEMMs: one launch
            self._adam_segment("heads")
            a_in = self.act[self.enc_chain[-1]]
            if train_z and train_c:
                self._wgrad("ench", a_in, a_in.stride(0), dh, dh.stride(0), rows)
            elif train_c:      # prior pre-training: only the c-head columns (base_models.py:312-321)
                self._wgrad("ench", a_in, a_in.stride(0), dh[:, hp:], dh.stride(0), rows, n_out=hp, col0=hp)
            elif train_z:
                self._wgrad("ench", a_in, a_in.stride(0), dh, dh.stride(0), rows, n_out=hp, col0=0)
            if train_trunk:
                lp = self.layers[self.enc_chain[-1]]
                if train_z and train_c:
                    self._dgrad("ench", dh, dh.stride(0), a_in, a_in.stride(0), self.dact[self.enc_chain[-1]], dt, rows,
                                lp.n_valid, lp.n_block)
                else:
                    c0 = hp if train_c else 0
                    Wsub = self.W("ench", op=True)[:, c0:]
                    self._dgrad("ench", dh[:, c0:], dh.stride(0), a_in, a_in.stride(0), self.dact[self.enc_chain[-1]], dt,
                                rows, lp.n_valid, lp.n_block, W=Wsub, ldw=self.layers["ench"].out_pad, n_out_pad=hp)
            self._flush_group()
            self._adam_segment("ench")
        else:
            if train_z:
                a_in = self.act[self.enc_chain[-1]]
                lp = self.layers[self.enc_chain[-1]]
                self._wgrad("zh", a_in, a_in.stride(0), self.dzh, self.dzh.stride(0), rows)
                self._dgrad("zh", self.dzh, self.dzh.stride(0), a_in, a_in.stride(0), self.dact[self.enc_chain[-1]], dt,
                            rows, lp.n_valid, lp.n_block)
                self._flush_group()
                self._adam_segment("heads")
        # ---- encoder trunk ----
        if train_trunk:
            ec = self.enc_chain
            for i in range(len(ec) - 1, -1, -1):
                nm = ec[i]
                a_in = self.act[ec[i - 1]] if i > 0 else self.act["x"]
                dy = self.dact[nm]
                self._wgrad(nm, a_in, a_in.stride(0), dy, dy.stride(0), rows)
                if i > 0:
                    lp = self.layers[ec[i - 1]]
                    self._dgrad(nm, dy, dy.stride(0), a_in, a_in.stride(0), self.dact[ec[i - 1]], dt, rows, lp.n_valid,
                                lp.n_block)
                self._flush_group()
                if i > 0:
                    self._adam_segment("enc:" + nm)
        self._end_group()
        self._join()

    # ------------------------------------------------------------------------------------------
    # optimiser
    # ------------------------------------------------------------------------------------------
    def optimizer(self, key: str, lr: float) -> AdamState:
        if key not in self.optimizers:
            self.optimizers[key] = AdamState(self.n_params, self.device, lr)
        return self.optimizers[key]

    def zero_grads(self):
        _abi.check(self.lib.dmvae_zero_f32(self.ctx, self.grads.data_ptr(), self.n_params, self._stream()))
        self._grads_dirty = False

    def adam(self, opt: AdamState, zero_grads: bool = True, use_dev: bool = False):
        lr_t = 0.0 if use_dev else opt.next_lr_t()
        t0 = self._tic("adam")
        _abi.check(self.lib.dmvae_adam(self.ctx, self.params.data_ptr(), self.grads.data_ptr(), opt.m.data_ptr(),
                                       opt.v.data_ptr(), self.params_op.data_ptr() if self.params_op is not None else None,
                                       self.n_params, lr_t, (opt.state_dev.data_ptr() + 12) if use_dev else None,
                                       opt.beta1, opt.beta2, opt.eps, 1.0, 1 if zero_grads else 0, self._stream()))
        self._toc("adam", t0)
        self._refresh_split_heads()
        if zero_grads:
            self._grads_dirty = False

    # ---- streamed update: Adam runs on a layer block as soon as that block's weight gradient is final (and the data
    #      gradient that reads the block's bf16 operand copy is done), on the side stream beside the remaining gradient
    #      GEMMs - those are bound by operand delivery from L2, the update by HBM, so the two overlap almost freely and
    #      only the first encoder layer's block is left for the tail of the step.  Captured (device-scalar) steps only.
    #      DMVAE_STREAM_ADAM=0 restores the single update at the end. ----
    stream_adam = os.environ.get("DMVAE_STREAM_ADAM", "1") != "0"
    _adam_live = None              # (AdamState, [(offset, n) already updated]) while a captured step streams its update

    def _adam_range(self, opt: AdamState, off: int, n: int, background: bool = False):
        _abi.check(self.lib.dmvae_adam(self.ctx, self.params.data_ptr() + 4 * off, self.grads.data_ptr() + 4 * off,
                                       opt.m.data_ptr() + 4 * off, opt.v.data_ptr() + 4 * off,
                                       (self.params_op.data_ptr() + 2 * off) if self.params_op is not None else None,
                                       n, 0.0, opt.state_dev.data_ptr() + 12, opt.beta1, opt.beta2, opt.eps, 1.0,
                                       3 if background else 1, self._stream()))

    def _block_range(self, name: str) -> Tuple[int, int]:
        return self.layout.block_range(name)

    def _stream_plan(self) -> Dict[str, List[str]]:
        return self.layout.stream_plan()

    def _merged_ranges(self, names) -> List[Tuple[int, int]]:
        return self.layout.merged_ranges(names)

    def stream_partition(self) -> List[Tuple[int, int]]:
        return self.layout.stream_partition()

    def _adam_segment(self, key: str):
        """The blocks of plan entry `key` are final: update them now on the side stream (contiguous blocks share one
        launch; with data parallelism: exchange + update of that range)."""
        if self._adam_live is None:
            return
        names = self._stream_plan().get(key)
        if not names:
            return
        if self.dp is not None and (key.startswith("enc") or key == "ench"):
            # measured (2 x B200 timelines): an exchange kernel that becomes runnable while a gradient GEMM is in full swing
            # only starts at that GEMM's end, so a segment that is final late in the backward pass would be exchanged
            # after the last GEMM anyway - plus two more cross-GPU barriers; those go with the final exchange
            return
        opt, done = self._adam_live
        merged = self._merged_ranges(names)

        def run():
            if self.dp is not None:
                self.dp.segment(opt, merged)
                return
            for off, n in merged:
                self._adam_range(opt, off, n, background=self.dt == BF16)
            if "ch" in names:
                self._refresh_split_heads()

        self._fork(run)
        done.extend(merged)

    def _adam_rest(self, opt: AdamState):
        """The blocks no segment covered (the tail of the step)."""
        done = sorted(self._adam_live[1])
        self._adam_live = None
        if self.dp is not None:
            self.dp.update(opt, True, skip=done)
            return
        pos = 0
        ch = self.layers.get("ch")
        for off, n in done + [(self.n_params, 0)]:
            if off > pos:
                self._adam_range(opt, pos, off - pos)
                if ch is not None and pos <= ch.offset < off:
                    self._refresh_split_heads()
            pos = max(pos, off + n)
        self._grads_dirty = False

    # ------------------------------------------------------------------------------------------
    # whole steps
    # ------------------------------------------------------------------------------------------
    def forward_backward(self, X: torch.Tensor, rows: int, eps: Optional[torch.Tensor] = None,
                         gumbel: Optional[torch.Tensor] = None, kl_ratio: float = 1.0, inv_global_batch=None,
                         row_offset: int = 0, recon_scale: float = 1.0, backward: bool = True, mode: str = "all",
                         dev_state: Optional[AdamState] = None, fuse: Optional[bool] = None):
        """encoder -> reparam -> decoder -> fused ELBO (-> gradient GEMMs).  Results stay on the device.
        dev_state: read the Philox step / kl_ratio from device memory (CUDA-graph capture)."""
        if dev_state is None:
            if rows > self.max_rows:
                self._alloc_activations(rows)
            if getattr(self, "_params_dirty", False):
                self.sync_operand_copy()
            if backward and getattr(self, "_grads_dirty", False):
                self.zero_grads()                  # split-K wgrads accumulate into the gradient buffer
            if eps is not None:
                self.eps_in[:rows].copy_(eps.reshape(rows, self.L))
            if gumbel is not None:
                self.gumbel_in[:rows].copy_(gumbel.reshape(rows, self.K))
        X, xdt = self.stage_input(X, rows)
        sdev = dev_state.state_dev.data_ptr() if dev_state is not None else None
        klr_dev = self.klr_dev.data_ptr() if (dev_state is not None and mode == "all") else None
        # fused reconstruction term: the captured training step only by default - decoded_X is not produced, so every
        # path whose results can be fetched (eager steps, evaluation) keeps the separate ELBO kernel
        if fuse is None:
            fuse = self.fuse_recon and backward and dev_state is not None
        fuse = fuse and self._fuse_ok(X, xdt, rows) and not self._chain_ok(gumbel is not None)
        fused = None
        if self._chain_ok(gumbel is not None):
            self.forward_chain(rows, eps is not None, row_offset, step_dev=sdev)
        else:
            self.encode(rows, fold_in_reparam=True)
            self.reparam(rows, eps is not None, gumbel is not None, row_offset, step_dev=sdev)
            if fuse:
                fused = self.elbo_latent_fork(X, xdt, rows, kl_ratio, inv_global_batch, recon_scale, klr_dev,
                                              prior_grads=(backward and mode == "all"))
                s = (1.0 / rows) if inv_global_batch is None else inv_global_batch
                self.decode(rows, recon=(X, xdt, s * recon_scale))
            else:
                self.decode(rows)
        flags = dict(all=(True, True, True, True), vae=(True, True, False, True), prior=(False, False, True, False))[mode]
        # prior-table gradients only when a full training step will consume them: an evaluation pass (backward=False) or
        # a pre-training mode must leave the means / log_vars gradient slots alone (base_models.py:307-321 never touch them)
        self.elbo(X, xdt, rows, kl_ratio, inv_global_batch, recon_scale, prior_grads=(backward and mode == "all"),
                  klr_dev=klr_dev, fused=fused)
        if backward:
            self.backward(rows, train_decoder=flags[0], train_z=flags[1], train_c=flags[2], train_trunk=flags[3])
            self._grads_dirty = True
        self._join()

    # ------------------------------------------------------------------------------------------
    # mixture-of-experts head (models.py:53-111, :149-163)
    # ------------------------------------------------------------------------------------------
    def moe_step(self, X: torch.Tensor, Y: torch.Tensor, rows: int, opt: Optional[AdamState], eps=None, gumbel=None,
                 kl_ratio: float = 1.0, train: bool = True, graph: bool = False):
        """Gate = q(c|x) of the VAE, experts = one dense GEMM over all experts; supervised loss (+ the VAE loss when
        lossVAE).  Fills moe_loss = [supervised loss, error] and loss_out (VAE terms when lossVAE).
        graph=True (persistent X / Y buffers, device noise): the step is captured once into a CUDA graph and replayed,
        with Adam's lr_t and the Philox step read from device memory."""
        if graph and train and opt is not None and eps is None and gumbel is None and self.use_graphs:
            return self._moe_step_graph(X, Y, rows, opt, kl_ratio)
        if rows > self.max_rows:
            self._alloc_activations(rows)
        if getattr(self, "_params_dirty", False):
            self.sync_operand_copy()
        if train and getattr(self, "_grads_dirty", False):
            self.zero_grads()
        if eps is not None:
            self.eps_in[:rows].copy_(eps.reshape(rows, self.L))
        if gumbel is not None:
            self.gumbel_in[:rows].copy_(gumbel.reshape(rows, self.K))
        self._moe_body(X, Y, rows, eps is not None, gumbel is not None, kl_ratio, train, None)
        if train and opt is not None:
            self._update(opt)
            self.step_count += 1

    def _moe_step_graph(self, X, Y, rows, opt, kl_ratio):
        key = ("moe", X.data_ptr(), X.dtype, Y.data_ptr(), rows, id(opt), float(kl_ratio), float(self.x_scale))
        ent = self._graphs.get(key)
        if ent is None:
            self._log_moe = (None, self.step_count)
            self.moe_step(X, Y, rows, opt, kl_ratio=kl_ratio, graph=False)       # eager: warms descriptor caches
            self._log_moe = None
            torch.cuda.current_stream(self.device).synchronize()
            g = torch.cuda.CUDAGraph()
            l0 = int(self.lib.dmvae_ctx_launch_count(self.ctx))
            with torch.cuda.graph(g):
                self._fork(lambda: _abi.check(self.lib.dmvae_step_tick(self.ctx, opt.state_dev.data_ptr(), opt.lr,
                                                                        opt.beta1, opt.beta2, self._stream())))
                self._log_moe = (opt.state_dev.data_ptr(), 0)
                self._moe_body(X, Y, rows, False, False, kl_ratio, True, opt)
                self._log_moe = None
                self._update(opt, use_dev=True)
            self._graphs[key] = (g, int(self.lib.dmvae_ctx_launch_count(self.ctx)) - l0, False)
            self._grads_dirty = False
            return
        g, n_nodes, _ = ent
        if getattr(self, "_params_dirty", False):
            self.sync_operand_copy()
        if getattr(self, "_grads_dirty", False):
            self.zero_grads()
        if not opt.state_valid:
            opt.upload_state(self.step_count)
        g.replay()
        self._grads_dirty = False
        opt.t += 1
        self.step_count += 1
        self._graph_replay_launches += n_nodes

    def _moe_body(self, X, Y, rows, eps_injected, gumbel_injected, kl_ratio, train, dev_state):
        cfg = self.moe
        E, O, lossVAE, feat = cfg["n_experts"], cfg["output_dim"], bool(cfg["lossVAE"]), bool(cfg["featLearn"])
        sdev = dev_state.state_dev.data_ptr() if dev_state is not None else None
        self.y_buf[:rows].copy_(Y.reshape(rows, O))
        st = self._stream
        Xs, xdt = self.stage_input(X, rows)
        full = lossVAE or self.model == "vade"
        if full:
            self.encode(rows, fold_in_reparam=True)
            self.reparam(rows, eps_injected, gumbel_injected, step_dev=sdev)
            self.decode(rows)
            self.elbo(Xs, xdt, rows, kl_ratio if lossVAE else 0.0, None, 1.0 if lossVAE else 0.0, prior_grads=lossVAE and train)
            self._join()
        else:
            self.encode(rows, heads=("z", "c") if feat else ("c",))
            self._join()
            _abi.check(self.lib.dmvae_softmax_rows(self.ctx, self.ch.data_ptr(), self.ch.stride(0), rows, self.K,
                                                   self.qc.data_ptr(), st()))
        # expert inputs
        if feat:
            _abi.check(self.lib.dmvae_stage_features(self.ctx, self.zh.data_ptr(), self.zh.stride(0), rows, self.L, 1,
                                                     self.moe_in.data_ptr(), self.dt, self.moe_in.stride(0),
                                                     self.moe_in.shape[1], st()))
            a_in = self.moe_in
        else:
            a_in = self.act["x"]
        self._fwd("moe", a_in, a_in.stride(0), self.moe_pred, F32, _abi.ACT_NONE, rows)
        ma = _abi.MoeArgs()
        ma.classification, ma.rows, ma.E, ma.O = int(bool(cfg["classification"])), rows, E, O
        ma.pred, ma.ld_pred = self.moe_pred.data_ptr(), self.moe_pred.stride(0)
        ma.gate, ma.ld_gate = self.qc.data_ptr(), self.K
        ma.Y, ma.ldy = self.y_buf.data_ptr(), O
        ma.inv_global_batch = 1.0 / rows
        ma.per_sample, ma.y_soft, ma.pred_class = self.moe_ps.data_ptr(), self.moe_ysoft.data_ptr(), self.moe_cls.data_ptr()
        ma.d_pred, ma.dpred_dtype = self.moe_dpred.data_ptr(), self.dt
        ma.ld_dpred, ma.dpred_cols = self.moe_dpred.stride(0), self.moe_dpred.shape[1]
        ma.d_gate, ma.ld_dgate = self.moe_dgate.data_ptr(), E
        _abi.check(self.lib.dmvae_moe_fwd_bwd(self.ctx, C.byref(ma), st()))
        _abi.check(self.lib.dmvae_reduce_columns(self.ctx, self.moe_ps.data_ptr(), 2, rows, 2, 1.0, self.moe_loss.data_ptr(), st()))
        if not train:
            return
        log = self._log_moe
        if log is not None:
            # both loss blocks are final here (the ELBO's reduction was joined above): file them in the rings on the side stream
            def append():
                _abi.check(self.lib.dmvae_log_append(self.ctx, self.moe_loss.data_ptr(), 2, self.moe_ring.data_ptr(),
                                                     self.LOSS_RING, log[0], log[1], self._stream()))
                _abi.check(self.lib.dmvae_log_append(self.ctx, self.loss_out.data_ptr(), 4, self.loss_ring.data_ptr(),
                                                     self.LOSS_RING, log[0], log[1], self._stream()))
            self._fork(append)
        if self.model == "vade":
            # the gate is gamma = get_cluster_probs(Z) (models.py:74, priors.py:91-102): the supervised loss reaches Z, mean,
            # log_var and the prior tables through it.  Second pass of the fused ELBO kernel with d loss / d gamma added
            # through gamma's softmax Jacobian (the first pass produced gamma for the experts' mixture).
            self._join()
            self.elbo(Xs, xdt, rows, kl_ratio if lossVAE else 0.0, None, 1.0 if lossVAE else 0.0, prior_grads=True,
                      d_gate_extra=self.moe_dgate)
            self._join()
            if not lossVAE:                      # no decoder backward: dZ has the gamma path only
                _abi.check(self.lib.dmvae_zero_f32(self.ctx, self.dz.data_ptr(), rows * self.dz.stride(0), st()))
        else:
            # gate gradient through the softmax into the c-head logits (added to the ELBO's own d_logits when lossVAE)
            _abi.check(self.lib.dmvae_softmax_bwd_add(self.ctx, rows, self.K, self.qc.data_ptr(), self.moe_dgate.data_ptr(), E,
                                                      self.dch.data_ptr(), self.dt, self.dch.stride(0), 1 if lossVAE else 0,
                                                      self.dch.shape[1], st()))
        self._wgrad("moe", a_in, a_in.stride(0), self.moe_dpred, self.moe_dpred.stride(0), rows)
        dme = None
        if feat:
            # d relu(mean): dgrad through the expert weights, masked by relu(mean) > 0, added to d_mean
            self._dgrad("moe", self.moe_dpred, self.moe_dpred.stride(0), self.moe_in, self.moe_in.stride(0), self.moe_dinp, F32,
                        rows, self.L, self.moe_dinp.shape[1])
            dme = self.moe_dinp
        self.backward(rows, dmean_extra=dme, train_decoder=lossVAE, train_z=lossVAE or feat, train_c=True, train_trunk=True,
                      through_decoder=lossVAE)
        self._grads_dirty = True
        self._join()

    use_graphs = True

    def train_step(self, X: torch.Tensor, rows: int, opt: AdamState, eps=None, gumbel=None, kl_ratio=1.0,
                   mode: str = "all", recon_scale: float = 1.0, graph: Optional[bool] = None):
        """One optimisation step (VAE.train_op body, base_models.py:117-129).  Returns nothing; read loss_out.
        Without injected noise the step is captured once per (input buffer, rows, mode, optimiser) into a CUDA graph
        and replayed: per-step scalars (Adam's lr_t, the Philox step, kl_ratio) live in device memory."""
        use_graph = self.use_graphs if graph is None else graph
        if use_graph and eps is None and gumbel is None:
            return self._train_step_graph(X, rows, opt, kl_ratio, mode, recon_scale)
        inv, off = self._dp_scale(rows)
        self._dp_opt = (opt, False) if mode == "all" else None
        self.forward_backward(X, rows, eps, gumbel, kl_ratio, inv, off, recon_scale, True, mode)
        self._dp_opt = None
        self._update(opt)
        self.step_count += 1
        if self.dp is not None:
            self.dp.mark_updated()

    def _dp_scale(self, rows):
        """(1 / global batch, global index of this rank's first row): data-parallel shards."""
        if self.world == 1:
            return None, 0
        return 1.0 / (rows * self.world), self.rank * rows

    def _update(self, opt, use_dev: bool = False):
        if self.dp is None:
            self.adam(opt, zero_grads=True, use_dev=use_dev)
        else:
            self.dp.update(opt, use_dev)

    def _train_step_graph(self, X, rows, opt, kl_ratio, mode, recon_scale):
        key = (X.data_ptr(), X.dtype, X.stride(0), rows, mode, id(opt), float(recon_scale),
               float(kl_ratio) if mode != "all" else None, float(self.x_scale))
        ent = self._graphs.get(key)
        if ent is None:
            # first use of this signature: one eager step (also warms the TMA-descriptor cache and the kernels'
            # shared-memory attributes), then capture the same sequence for every later step
            inv, off = self._dp_scale(rows)
            self._dp_opt = (opt, False) if mode == "all" else None
            self._log_step = (None, self.step_count)
            self.forward_backward(X, rows, None, None, kl_ratio, inv, off, recon_scale, True, mode)
            self._dp_opt = None
            self._log_step = None
            self._update(opt)
            self.step_count += 1
            torch.cuda.current_stream(self.device).synchronize()
            g = torch.cuda.CUDAGraph()
            l0 = int(self.lib.dmvae_ctx_launch_count(self.ctx))
            with torch.cuda.graph(g):
                # the per-step scalars (Philox step, Adam lr_t) are first read by the reparameterisation: advance them on
                # the side stream, off the head of the critical path
                self._fork(lambda: _abi.check(self.lib.dmvae_step_tick(self.ctx, opt.state_dev.data_ptr(), opt.lr,
                                                                        opt.beta1, opt.beta2, self._stream())))
                # data parallel (peer-memory exchange): the gradients consumed by the previous step's exchange are cleared
                # here, beside the forward pass, rather than as the last kernel of that step
                defer = self.dp is not None and getattr(self.dp, "mode", "") == "p2p" and mode == "all"
                if defer:
                    self._fork(lambda: _abi.check(self.lib.dmvae_zero_f32(self.ctx, self.grads.data_ptr(), self.n_params,
                                                                          self._stream())))
                    self.dp.defer_clear = True
                self._dp_opt = (opt, True) if mode == "all" else None
                if (self.stream_adam and mode == "all" and self.timers is None and self.overlap
                        and (self.dp is None or self.dp.can_stream())):
                    self._adam_live = (opt, [])
                self._log_step = (opt.state_dev.data_ptr(), 0)
                self.forward_backward(X, rows, None, None, kl_ratio, inv, off, recon_scale, True, mode, dev_state=opt)
                self._log_step = None
                self._dp_opt = None
                if self._adam_live is not None:
                    self._adam_rest(opt)
                else:
                    self._update(opt, use_dev=True)
                if defer:
                    self.dp.defer_clear = False
            n_nodes = int(self.lib.dmvae_ctx_launch_count(self.ctx)) - l0
            self._graphs[key] = (g, n_nodes, defer)
            self._grads_dirty = False                  # the eager step above cleared them (the capture launched nothing)
            if self.dp is not None:
                self.dp.mark_updated()
            return
        g, n_nodes, self_clearing = ent
        if getattr(self, "_params_dirty", False):
            self.sync_operand_copy()
        if getattr(self, "_grads_dirty", False) and not self_clearing:
            self.zero_grads()
        if not opt.state_valid:
            opt.upload_state(self.step_count)
        if mode == "all" and kl_ratio != self._klr_host:
            self.klr_dev.fill_(float(kl_ratio))
            self._klr_host = float(kl_ratio)
        g.replay()
        self._grads_dirty = self_clearing              # a self-clearing graph leaves the consumed gradients for its next replay
        if self.dp is not None:
            self.dp.mark_updated()
        opt.t += 1                     # host mirrors of the device counters
        self.step_count += 1
        self._graph_replay_launches += n_nodes

    def run_epoch(self, host: torch.Tensor, batch_size: int, opt: AdamState, kl_ratio: float = 1.0, mode: str = "all",
                  max_steps: Optional[int] = None, perm: Optional[np.ndarray] = None, while_busy=None,
                  x_scale: float = 1.0, packed_D: int = 0) -> float:
        """One pass over a (pinned) host array [N, D] - or, with ``packed_D`` = D, over {0,1}-valued rows stored one bit per
        element ([N, pitch] uint8 from includes/utils.py::_pack_bits), which the gather expands (dmvae_gather_rows_bits).  Per step, on a copy stream and double-buffered so that it overlaps
        the previous step: the batch's rows cross the bus into a staging buffer - a plain asynchronous copy of a
        contiguous slice, or with ``perm`` (the epoch's shuffle, includes/utils.py:450-454) a gather kernel that reads the
        rows ``perm[lo:lo+B]`` straight from the pinned host array (dmvae_gather_rows) - then the training step, and an
        asynchronous read of its loss.  One synchronisation at the end.  Returns the mean batch loss
        (base_models.py:130)."""
        N = host.shape[0]
        self.x_scale = float(x_scale)
        nb = (N + batch_size - 1) // batch_size
        if max_steps is not None:
            nb = min(nb, max_steps)
        dev = self.device
        if batch_size > self.max_rows:
            self._alloc_activations(batch_size)
        if packed_D:
            assert packed_D == self.D and host.dtype == torch.uint8
            if perm is None:
                perm = np.arange(N, dtype=np.int32)            # the expanding gather is index-driven
        key = (batch_size, host.dtype)
        if getattr(self, "_epoch_key", None) != key:
            self._epoch_key = key
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._stage = [torch.empty(batch_size, self.D, dtype=host.dtype, device=dev) for _ in range(2)]
            self._ready = [torch.cuda.Event() for _ in range(2)]
            self._free = [torch.cuda.Event() for _ in range(2)]
        if getattr(self, "_loss_log", None) is None or self._loss_log.shape[0] < nb:
            self._loss_log = torch.zeros(nb, 4, dtype=torch.float32, device=dev)
            self._loss_host = torch.zeros(nb, 4, dtype=torch.float32).pin_memory()
        cur = torch.cuda.current_stream(dev)
        perm_dev = None
        if perm is not None:
            if not host.is_pinned():
                raise ValueError("run_epoch(perm=...) gathers from the host array on the device: it must be pinned")
            if getattr(self, "_perm_dev", None) is None or self._perm_dev.numel() < N:
                self._perm_dev = torch.empty(N, dtype=torch.int32, device=dev)
                self._perm_pin = torch.empty(N, dtype=torch.int32).pin_memory()
            cur.synchronize()                      # the previous epoch's gathers have consumed the old permutation
            self._perm_pin[:N].copy_(torch.from_numpy(np.ascontiguousarray(perm[:N], dtype=np.int32)))
            with torch.cuda.stream(self._copy_stream):
                self._perm_dev[:N].copy_(self._perm_pin[:N], non_blocking=True)
            perm_dev = self._perm_dev
        row_bytes = host.shape[1] * host.element_size()
        ring = nb <= self.LOSS_RING and self.use_graphs and os.environ.get("DMVAE_LOSS_RING", "1") != "0"
        slots = []
        klr, rs = kl_ratio, 1.0
        if mode == "vae":
            klr = 0.0
        elif mode == "prior":
            klr, rs = 1.0, 0.0
        for i in range(nb):
            b = i & 1
            lo = i * batch_size
            rows = min(batch_size, N - lo)
            with torch.cuda.stream(self._copy_stream):
                if i >= 2:
                    self._copy_stream.wait_event(self._free[b])
                if packed_D:
                    _abi.check(self.lib.dmvae_gather_rows_bits(self.ctx, host.data_ptr(), host.stride(0),
                                                               perm_dev.data_ptr() + 4 * lo, self._stage[b].data_ptr(),
                                                               self._stage[b].stride(0), rows, packed_D,
                                                               C.c_void_p(self._copy_stream.cuda_stream)))
                elif perm_dev is None:
                    self._stage[b][:rows].copy_(host[lo:lo + rows], non_blocking=True)
                else:
                    _abi.check(self.lib.dmvae_gather_rows(self.ctx, host.data_ptr(), host.stride(0) * host.element_size(),
                                                          perm_dev.data_ptr() + 4 * lo, self._stage[b].data_ptr(), row_bytes,
                                                          rows, row_bytes, C.c_void_p(self._copy_stream.cuda_stream)))
                self._ready[b].record(self._copy_stream)
            cur.wait_event(self._ready[b])
            if ring and not slots:
                slots.append(self.step_count % self.LOSS_RING)
            self.train_step(self._stage[b], rows, opt, None, None, klr, mode, rs)
            if not ring:
                self._loss_log[i].copy_(self.loss_out, non_blocking=True)
            self._free[b].record(cur)
        if not ring:
            self._loss_host[:nb].copy_(self._loss_log[:nb], non_blocking=True)
        if while_busy is not None:
            while_busy()                           # host work hidden behind the queued steps (e.g. the next epoch's shuffle)
        cur.synchronize()
        col = {"all": 3, "vae": 0, "prior": 3}[mode]
        if ring:
            # the steps wrote their loss terms into the pinned ring themselves (slot = step counter): just read them
            idx = (slots[0] + np.arange(nb)) % self.LOSS_RING
            return float(self.loss_ring.numpy()[idx, col].astype(np.float64).sum()) / nb
        return float(self._loss_host[:nb, col].sum()) / nb

    def run_epoch_moe(self, host_x: torch.Tensor, host_y: torch.Tensor, batch_size: int, opt: AdamState, kl_ratio: float = 1.0,
                      perm: Optional[np.ndarray] = None, max_steps: Optional[int] = None, while_busy=None,
                      x_scale: float = 1.0, packed_D: int = 0) -> np.ndarray:
        """One pass of the MoE training step (models.py:194-221) over pinned host arrays X [N, D] / Y [N, O], batches staged
        like run_epoch.  Returns per-step [supervised loss sum, error sum, recon, KL_c, KL_z, VAE loss] as a host array."""
        N = host_x.shape[0]
        self.x_scale = float(x_scale)
        O = host_y.shape[1]
        nb = (N + batch_size - 1) // batch_size
        if max_steps is not None:
            nb = min(nb, max_steps)
        dev = self.device
        if batch_size > self.max_rows:
            self._alloc_activations(batch_size)
        if packed_D:
            assert packed_D == self.D and host_x.dtype == torch.uint8
            if perm is None:
                perm = np.arange(N, dtype=np.int32)
        key = ("moe", batch_size, host_x.dtype, O)
        if getattr(self, "_epoch_key", None) != key:
            self._epoch_key = key
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._stage = [torch.empty(batch_size, self.D, dtype=host_x.dtype, device=dev) for _ in range(2)]
            self._stage_y = [torch.empty(batch_size, O, dtype=torch.float32, device=dev) for _ in range(2)]
            self._ready = [torch.cuda.Event() for _ in range(2)]
            self._free = [torch.cuda.Event() for _ in range(2)]
        if getattr(self, "_moe_log", None) is None or self._moe_log.shape[0] < nb:
            self._moe_log = torch.zeros(nb, 6, dtype=torch.float32, device=dev)
            self._moe_log_host = torch.zeros(nb, 6, dtype=torch.float32).pin_memory()
        cur = torch.cuda.current_stream(dev)
        perm_dev = None
        if perm is not None:
            if getattr(self, "_perm_dev", None) is None or self._perm_dev.numel() < N:
                self._perm_dev = torch.empty(N, dtype=torch.int32, device=dev)
                self._perm_pin = torch.empty(N, dtype=torch.int32).pin_memory()
            cur.synchronize()
            self._perm_pin[:N].copy_(torch.from_numpy(np.ascontiguousarray(perm[:N], dtype=np.int32)))
            with torch.cuda.stream(self._copy_stream):
                self._perm_dev[:N].copy_(self._perm_pin[:N], non_blocking=True)
            perm_dev = self._perm_dev
        xb, yb = host_x.shape[1] * host_x.element_size(), O * 4
        cs = C.c_void_p(self._copy_stream.cuda_stream)
        ring = nb <= self.LOSS_RING and self.use_graphs and os.environ.get("DMVAE_LOSS_RING", "1") != "0"
        first = 0
        for i in range(nb):
            b = i & 1
            lo = i * batch_size
            rows = min(batch_size, N - lo)
            with torch.cuda.stream(self._copy_stream):
                if i >= 2:
                    self._copy_stream.wait_event(self._free[b])
                if perm_dev is None:
                    self._stage[b][:rows].copy_(host_x[lo:lo + rows], non_blocking=True)
                    self._stage_y[b][:rows].copy_(host_y[lo:lo + rows], non_blocking=True)
                else:
                    ix = perm_dev.data_ptr() + 4 * lo
                    if packed_D:
                        _abi.check(self.lib.dmvae_gather_rows_bits(self.ctx, host_x.data_ptr(), host_x.stride(0), ix,
                                                                   self._stage[b].data_ptr(), self._stage[b].stride(0), rows,
                                                                   packed_D, cs))
                    else:
                        _abi.check(self.lib.dmvae_gather_rows(self.ctx, host_x.data_ptr(), xb, ix, self._stage[b].data_ptr(), xb, rows, xb, cs))
                    _abi.check(self.lib.dmvae_gather_rows(self.ctx, host_y.data_ptr(), yb, ix, self._stage_y[b].data_ptr(), yb, rows, yb, cs))
                self._ready[b].record(self._copy_stream)
            cur.wait_event(self._ready[b])
            if i == 0:
                first = self.step_count % self.LOSS_RING
            self.moe_step(self._stage[b], self._stage_y[b], rows, opt, kl_ratio=kl_ratio, graph=True)
            if not ring:
                self._moe_log[i, :2].copy_(self.moe_loss, non_blocking=True)
                self._moe_log[i, 2:].copy_(self.loss_out, non_blocking=True)
            self._free[b].record(cur)
        if not ring:
            self._moe_log_host[:nb].copy_(self._moe_log[:nb], non_blocking=True)
        if while_busy is not None:
            while_busy()
        cur.synchronize()
        if ring:
            # the captured steps wrote their loss blocks into the pinned rings (slot = step counter)
            idx = (first + np.arange(nb)) % self.LOSS_RING
            return np.concatenate([self.moe_ring.numpy()[idx], self.loss_ring.numpy()[idx]], axis=1)
        return self._moe_log_host[:nb].numpy().copy()

    def launches(self) -> int:
        """Kernels launched so far: direct launches counted by the library + nodes of replayed CUDA graphs."""
        return int(self.lib.dmvae_ctx_launch_count(self.ctx)) + getattr(self, "_graph_replay_launches", 0)

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.dmvae_ctx_destroy(self.ctx)
            self.ctx = None
