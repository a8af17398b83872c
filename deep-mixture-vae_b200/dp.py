"""Data parallelism for the training step (new functionality: the reference is single-device).

One process per GPU (torchrun); ``torch.distributed`` is the plumbing.  The minibatch is sharded by rank, every
rank scales its gradients by 1/global_batch inside the fused ELBO kernel, and the Philox counters are offset by
the global row index, so an N-GPU step computes exactly the single-GPU step of the concatenated batch.

The only exchange is the sum of the flat fp32 gradient buffer, fused with the Adam update:

  mode "p2p"  : the flat parameter / gradient buffers live in symmetric (peer-mapped) memory; ONE kernel per rank
                (dmvae_dp_reduce_adam) reads its 1/N shard of every peer's gradients over NVLink, sums them in a
                fixed order, applies Adam to the shard and stores the new fp32 + bf16 parameters into every
                replica (reduce-scatter + Adam + all-gather without a round trip through HBM).
  mode "nccl" : ncclAllReduce of the flat gradient buffer, then the replicated flat Adam kernel (baseline).

The host-side arithmetic (shard ranges, bucket layout) is device-agnostic and covered by gloo tests on CPU.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int, align: int = 4) -> Tuple[int, int]:
    """[begin, end) of rank's shard of n elements; boundaries are multiples of ``align`` (float4 access)."""
    per = (n + world - 1) // world
    per = (per + align - 1) // align * align
    b = min(n, rank * per)
    e = min(n, b + per)
    return b, e


def global_row_offset(rank: int, rows_per_rank: int) -> int:
    """Global index of a rank's first sample: the Philox counter offset that makes an N-GPU run reproduce the
    noise of the 1-GPU run on the concatenated batch."""
    return rank * rows_per_rank


def allreduce_adam_reference(grads: torch.Tensor, apply_adam, group=None):
    """Device-agnostic statement of mode "nccl": sum the gradients over ranks, then the same Adam everywhere."""
    dist.all_reduce(grads, op=dist.ReduceOp.SUM, group=group)
    apply_adam(0, grads.numel(), grads)


def sharded_adam_reference(grads: torch.Tensor, params: torch.Tensor, apply_adam, rank: int, world: int, group=None):
    """Device-agnostic statement of mode "p2p" with collectives: reduce-scatter the gradient shards, Adam on the owned
    shard, all-gather the parameters.  Must give the same parameters as allreduce_adam_reference."""
    n = grads.numel()
    per = shard_range(n, 0, world)[1]
    padded = torch.zeros(per * world, dtype=grads.dtype, device=grads.device)
    padded[:n] = grads
    mine = torch.zeros(per, dtype=grads.dtype, device=grads.device)
    parts = list(padded.view(world, per).unbind(0))
    dist.reduce_scatter(mine, [p.contiguous() for p in parts], op=dist.ReduceOp.SUM, group=group) \
        if dist.get_backend(group) != "gloo" else _gloo_reduce_scatter(mine, parts, rank, group)
    b, e = shard_range(n, rank, world)
    apply_adam(b, e, mine[: e - b])
    pp = torch.zeros(per * world, dtype=params.dtype, device=params.device)
    pp[:n] = params
    outs = [torch.zeros(per, dtype=params.dtype, device=params.device) for _ in range(world)]
    dist.all_gather(outs, pp.view(world, per)[rank].contiguous(), group=group)
    params.copy_(torch.cat(outs)[:n])


def _gloo_reduce_scatter(out, parts, rank, group):
    # gloo has no reduce_scatter: all-reduce every part and keep ours
    for r, p in enumerate(parts):
        t = p.clone()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        if r == rank:
            out.copy_(t)


class DataParallel:
    """Binds an Engine to the default process group."""

    def __init__(self, engine, mode: str = "auto", group=None):
        from . import _abi
        self._abi = _abi
        self.eng = engine
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        engine.world, engine.rank = self.world, self.rank
        engine.dp = self
        self.mode = mode
        self.hdl = None
        self.master_sharded = False
        self._master_stale = False
        if mode in ("auto", "p2p"):
            try:
                self._setup_symmetric()
                self.mode = "p2p"
            except Exception as ex:                      # no peer access / symmetric memory unavailable
                if mode == "p2p":
                    raise
                self.mode = "nccl"
                self.fallback_reason = repr(ex)
        # identical parameters everywhere (rank 0's initialisation)
        dist.broadcast(engine.params, src=0, group=group)
        engine.sync_operand_copy()
        torch.cuda.synchronize(engine.device)
        self._opt_shards = {}

    # ---- symmetric memory -------------------------------------------------------------------------------
    def _setup_symmetric(self):
        import torch.distributed._symmetric_memory as symm_mem
        eng = self.eng
        P = eng.n_params
        nbytes = 4 * P + 4 * P + (2 * P if eng.params_op is not None else 0)
        pad_off = (nbytes + 255) // 256 * 256             # flag pads of the library's own barrier (dmvae_dp_barrier)
        nbytes = pad_off + 4 * 8 * self.N_CHANNELS
        buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=eng.device)
        buf[pad_off:].zero_()
        grp = self.group if self.group is not None else dist.group.WORLD
        hdl = symm_mem.rendezvous(buf, grp)
        params = buf[: 4 * P].view(torch.float32)
        grads = buf[4 * P: 8 * P].view(torch.float32)
        params.copy_(eng.params)
        grads.zero_()
        eng.params, eng.grads = params, grads
        if eng.params_op is not None:
            pop = buf[8 * P: 10 * P].view(torch.bfloat16)
            pop.copy_(eng.params_op)
            eng.params_op = pop
        eng._graphs = {}                                  # captured graphs hold the old pointers
        if hasattr(eng, "_graph_replay_launches"):
            pass
        self.hdl, self._symm_buf = hdl, buf
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        VP = C.c_void_p * self.world
        self._g_ptrs = VP(*[p + 4 * P for p in ptrs])
        # bf16 tier: the GEMMs only read the bf16 operand copy, so only that is written into every replica; the fp32
        # master of a shard stays with its owner (Adam's only reader) and is fetched on demand (gather_master)
        self.master_sharded = eng.params_op is not None
        self._p_ptrs = VP(*[(p if (r == self.rank or not self.master_sharded) else None) for r, p in enumerate(ptrs)])
        self._p_ptrs_full = VP(*ptrs)                     # every replica's fp32 master (prior tables: read by the ELBO kernel)
        rr = eng.layout.replicated_master_ranges()
        self._n_rep = len(rr)
        self._rep_ranges = (C.c_int64 * (2 * len(rr)))(*[x for lo_hi in rr for x in lo_hi])
        self._b_ptrs = VP(*[(p + 8 * P) if eng.params_op is not None else 0 for p in ptrs])
        self._pad_ptrs = VP(*[p + pad_off for p in ptrs])
        # NVSwitch multicast mapping of the same buffer (0 when the fabric / driver does not offer it)
        self._mc = int(getattr(hdl, "multicast_ptr", 0) or 0) if self.use_multicast else 0
        self._mc_off = (4 * P, 8 * P)
        self._epochs = torch.zeros(self.N_CHANNELS, dtype=torch.int32, device=eng.device)
        torch.cuda.synchronize(eng.device)
        hdl.barrier(channel=0)                            # every pad is zeroed before anyone signals into it
        torch.cuda.synchronize(eng.device)
        self._master_stale = False

    N_CHANNELS = 32
    # DMVAE_DP_MULTICAST=1: in-switch reduction (multimem.ld_reduce) + broadcast store (multimem.st) through the NVSwitch
    # multicast mapping of the symmetric buffer instead of the peer-pointer kernel.  Correct (scripts/dp_check.py passes
    # with it) but measured SLOWER at N=2 (0.348 vs 0.331 ms / step): every GPU still serves its whole gradient buffer to
    # the switch (outbound bytes are unchanged, only inbound shrinks), the switch adds latency, and at N=2 the local
    # half of the sum also crosses NVLink.  Off by default; the peer-pointer kernel also keeps the rank-order sum.
    use_multicast = os.environ.get("DMVAE_DP_MULTICAST", "0") == "1"
    # DMVAE_DP_BARRIER=own: the library's own barrier kernel (dmvae_dp_barrier, flag pads in the symmetric buffer)
    # instead of torch's symmetric-memory barrier.  Same step time at N=2 (0.3312 ms either way); torch's is the default.
    own_barrier = os.environ.get("DMVAE_DP_BARRIER", "torch") == "own"

    def _barrier(self, ch: int):
        if self.own_barrier:
            eng = self.eng
            self._abi.check(eng.lib.dmvae_dp_barrier(eng.ctx, self.rank, self.world, self._pad_ptrs, self._epochs.data_ptr(),
                                                     ch, eng._stream()))
        else:
            self.hdl.barrier(channel=ch)

    # ---- ranges: the flat buffer is exchanged as encoder | decoder | tail (experts, prior tables).  The decoder's
    #      gradients are complete half way through the backward pass, so their exchange overlaps the encoder's
    #      backward GEMMs on a forked stream; the rest is exchanged at the end of the step. ----
    # Streamed exchange (default in mode "p2p"): the engine's backward pass hands over each range of the flat buffer as
    # soon as its gradients are final (Engine._adam_segment); barrier -> reduce + Adam on the owned shard -> barrier ->
    # local gradient clear run on the side stream beside the remaining gradient GEMMs, in 4-warp blocks that fit next to
    # the GEMM CTAs (see adam_bg_kernel).  Only the first encoder layer's range is left for the end of the step.
    # EXPERIMENTAL, off by default.  Measured on 2 x B200 (scripts/step_timeline.py under torchrun): an exchange kernel
    # that becomes runnable while a gradient GEMM is in full swing - which the opening barrier guarantees - only starts
    # when that GEMM kernel ends (the same kernel shape does run beside the GEMMs when it becomes runnable exactly at a
    # kernel boundary, as the N=1 streamed update does), so every streamed segment is late by one GEMM group and pays
    # two extra cross-GPU barriers: 0.337 ms / step with the two early segments streamed, 0.357 with three, 0.368 with
    # all four, against 0.331 for the single exchange.  scripts/dp_check.py's fp32 + CUDA-graph configuration did not
    # terminate with it.  DMVAE_DP_STREAM=1 enables it (bf16 tier).
    stream = os.environ.get("DMVAE_DP_STREAM", "0") == "1"

    def can_stream(self) -> bool:
        return (self.mode == "p2p" and self.stream and not self.overlap_decoder
                and getattr(self.eng, "params_op", None) is not None)

    def ranges(self):
        eng = self.eng
        if self.can_stream():
            return eng.stream_partition()
        if not self.overlap_decoder:
            return [(0, eng.n_params)]
        dec0 = eng.layers[eng.dec_chain[0]].offset
        dx = eng.layers["decx"]
        dec1 = dx.offset + dx.size
        return [(0, dec0), (dec0, dec1), (dec1, eng.n_params)]

    def range_shard(self, ridx, rank=None):
        lo, hi = self.ranges()[ridx]
        b, e = shard_range(hi - lo, self.rank if rank is None else rank, self.world)
        return lo + b, lo + e

    def _opt_shard_state(self, opt, ridx):
        """Adam slots of the owned shard of one range (the other ranks own the rest)."""
        key = (id(opt), ridx)
        if key not in self._opt_shards:
            b, e = self.range_shard(ridx)
            n = max(e - b, 4)
            self._opt_shards[key] = (torch.zeros(n, dtype=torch.float32, device=self.eng.device),
                                     torch.zeros(n, dtype=torch.float32, device=self.eng.device))
        return self._opt_shards[key]

    def segment(self, opt, merged):
        """Exchange + update of the partition ranges `merged` [(offset, n)] (captured steps: device-resident lr_t)."""
        rs = self.ranges()
        for off, n in merged:
            ridx = rs.index((off, off + n))
            self._exchange(opt, 0.0, opt.state_dev.data_ptr() + 12, [ridx], 2 + 2 * ridx, background=True)

    # Captured steps leave the consumed gradients in place and clear them at the START of the next replay, on the forked
    # stream beside the forward pass (Engine._train_step_graph), instead of as the last kernel of the step.
    defer_clear = False

    def _exchange(self, opt, lr_t, lr_dev, ridxs, ch, background=False):
        eng, abi = self.eng, self._abi
        self._barrier(ch)                                 # every rank's gradients of these ranges are complete
        for ridx in ridxs:
            b, e = self.range_shard(ridx)
            if e <= b:
                continue
            m, v = self._opt_shard_state(opt, ridx)
            # Some ranges need their fp32 master on EVERY rank even when the rest of the master is kept by its owner only
            # (Layout.replicated_master_ranges: the prior tables, read in fp32 by the fused ELBO kernel, and the logits
            # layer, whose split bf16 operand copy is rebuilt from the master): the kernel writes the fp32 values of those
            # ranges into every replica (one launch for the whole shard).
            if getattr(self, "_mc", 0) and not background:
                # in-switch reduction + broadcast store (NVLS): the multicast fp32 store is per launch, so the shard is
                # cut at the replicated ranges
                pieces, pos = [], b
                for lo, hi in (eng.layout.replicated_master_ranges() if self.master_sharded else []):
                    lo, hi = max(lo, b), min(hi, e)
                    if hi <= lo:
                        continue
                    if lo > pos:
                        pieces.append((pos, lo, False))
                    pieces.append((lo, hi, True))
                    pos = hi
                if pos < e:
                    pieces.append((pos, e, False))
                mc = self._mc
                for pb, pe, full in pieces:
                    abi.check(eng.lib.dmvae_dp_reduce_adam_mc(
                        eng.ctx, mc + self._mc_off[0], mc if (full or not self.master_sharded) else None,
                        (mc + self._mc_off[1]) if eng.params_op is not None else None, eng.params.data_ptr(),
                        m.data_ptr() + 4 * (pb - b), v.data_ptr() + 4 * (pb - b), eng.n_params, pb, pe, lr_t, lr_dev,
                        opt.beta1, opt.beta2, opt.eps, eng._stream()))
                continue
            abi.check(eng.lib.dmvae_dp_reduce_adam(eng.ctx, self.rank, self.world, self._g_ptrs, self._p_ptrs, self._b_ptrs,
                                                   self._p_ptrs_full if self.master_sharded else None,
                                                   self._rep_ranges if self.master_sharded else None,
                                                   self._n_rep if self.master_sharded else 0, m.data_ptr(), v.data_ptr(),
                                                   eng.n_params, b, e, lr_t, lr_dev, opt.beta1, opt.beta2, opt.eps,
                                                   2 if background else 0, eng._stream()))
        self._barrier(ch + 1)                             # every replica updated, every gradient shard consumed
        eng._refresh_split_heads()                        # logits layer: hi | lo | lo2 operand copy from the (replicated) master
        # clear the local gradients of these ranges (split-K accumulates into them): local HBM, not 7/8 remote stores
        if self.defer_clear:
            return
        for ridx in ridxs:
            lo, hi = self.ranges()[ridx]
            if hi > lo:
                abi.check(eng.lib.dmvae_zero_f32(eng.ctx, eng.grads.data_ptr() + 4 * lo, hi - lo, eng._stream()))

    # Measured on 2 x B200: the extra barrier pair and launches of the early exchange cost more than the overlap hides
    # (0.3645 vs 0.3558 ms / step), so one exchange at the end of the step is the default; DMVAE_DP_OVERLAP=1 enables it.
    overlap_decoder = os.environ.get("DMVAE_DP_OVERLAP", "0") == "1"

    def early(self, opt, use_dev: bool):
        """Exchange + Adam of the decoder range; called (on a forked stream) as soon as the decoder's weight gradients are
        queued.  update() then handles the other ranges."""
        if self.mode != "p2p" or not self.overlap_decoder:
            return
        self._lr_t_step = 0.0 if use_dev else opt.next_lr_t()
        lr_dev = (opt.state_dev.data_ptr() + 12) if use_dev else None
        self._exchange(opt, self._lr_t_step, lr_dev, [1], 2)
        self._early_done = True

    # ---- the exchange + update ---------------------------------------------------------------------------
    def update(self, opt, use_dev: bool = False, skip=None):
        """skip: [(offset, n)] ranges already exchanged by segment() during this step."""
        eng, abi = self.eng, self._abi
        if skip:
            lr_dev = opt.state_dev.data_ptr() + 12
            rs = self.ranges()
            gone = set((off, off + n) for off, n in skip)
            self._exchange(opt, 0.0, lr_dev, [i for i, r in enumerate(rs) if r not in gone], 0)
            self._master_stale = self.master_sharded
            eng._grads_dirty = False
            return
        early = getattr(self, "_early_done", False)
        self._early_done = False
        if early:
            lr_t = self._lr_t_step
        else:
            lr_t = 0.0 if use_dev else opt.next_lr_t()
        lr_dev = (opt.state_dev.data_ptr() + 12) if use_dev else None
        if self.mode == "p2p":
            self._exchange(opt, lr_t, lr_dev, [0, 2] if early else list(range(len(self.ranges()))), 0)
            self._master_stale = self.master_sharded
        else:
            dist.all_reduce(eng.grads, op=dist.ReduceOp.SUM, group=self.group)
            abi.check(eng.lib.dmvae_adam(eng.ctx, eng.params.data_ptr(), eng.grads.data_ptr(), opt.m.data_ptr(),
                                         opt.v.data_ptr(), eng.params_op.data_ptr() if eng.params_op is not None else None,
                                         eng.n_params, lr_t, lr_dev, opt.beta1, opt.beta2, opt.eps, 1.0, 1, eng._stream()))
            eng._refresh_split_heads()
        eng._grads_dirty = False

    def mark_updated(self):
        """Called once per optimisation step (also for CUDA-graph replays, where update() itself does not run)."""
        self._master_stale = self.master_sharded

    def gather_master(self):
        """Fetch the other ranks' fp32 master shards into this rank's parameter buffer (one-sided reads of the peers'
        symmetric memory; no collective, so a single rank may call it - e.g. rank 0 writing a checkpoint - as long as the
        ranks are between steps)."""
        if not getattr(self, "_master_stale", False) or self.hdl is None:
            return
        eng = self.eng
        torch.cuda.current_stream(eng.device).synchronize()
        n = eng.n_params
        for r in range(self.world):
            if r == self.rank:
                continue
            peer = self.hdl.get_buffer(r, (n,), torch.float32, 0)
            for ridx in range(len(self.ranges())):
                b, e = self.range_shard(ridx, r)
                if e > b:
                    eng.params[b:e].copy_(peer[b:e])
        torch.cuda.current_stream(eng.device).synchronize()
        self._master_stale = False

    def train_step(self, X, rows, opt, kl_ratio=1.0):
        self.eng.train_step(X, rows, opt, kl_ratio=kl_ratio)

    def run_epoch(self, host, batch_size, opt, kl_ratio=1.0, mode="all", max_steps=None, perm=None, while_busy=None,
                  x_scale=1.0, packed_D=0):
        """Every rank iterates over ITS host shard; the returned loss is the global mean (one all-reduce per epoch)."""
        loss = self.eng.run_epoch(host, batch_size, opt, kl_ratio, mode, max_steps, perm, while_busy, x_scale, packed_D)
        t = torch.tensor([loss], dtype=torch.float64, device=self.eng.device)
        dist.all_reduce(t, group=self.group)
        return float(t[0])
