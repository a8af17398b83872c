"""World-size-2 gloo tests (CPU) of the data-parallel host logic: shard arithmetic, and that the sharded
(reduce-scatter + Adam on the owned shard + all-gather) update equals the all-reduce + replicated Adam update and the
single-process update on the concatenated batch."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_shard_ranges_cover_and_align():
    from dmvae_b200.dp import shard_range, global_row_offset
    for n in (4, 64, 5312, 4373014 // 4 * 4, 1000):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(world):
                b, e = shard_range(n, r, world)
                assert b == prev and b % 4 == 0 and (e % 4 == 0 or e == n) and b <= e <= n
                prev = e
            assert prev == n
    assert global_row_offset(3, 4096) == 12288


def test_range_shards_partition_the_flat_buffer():
    """DataParallel.range_shard: with or without the early decoder exchange, the ranks' shards of all ranges tile the
    flat parameter buffer exactly once and stay float4-aligned (host logic only: a stub engine with the real Layout)."""
    from dmvae_b200 import dp
    from dmvae_b200.engine import Layout

    class Stub:
        pass
    lay = Layout(model="dmvae", input_dim=784, latent_dim=10, n_classes=10, trunk=(500, 500), head=2000,
                 decoder=(2000, 500, 500), name="dmvae")
    for overlap, stream in ((False, False), (True, False), (False, True)):
        for world in (1, 2, 4, 8):
            covered = np.zeros(lay.n_params, np.int32)
            for rank in range(world):
                d = dp.DataParallel.__new__(dp.DataParallel)
                eng = Stub()
                eng.layers, eng.dec_chain, eng.n_params = lay.layers, lay.dec_chain, lay.n_params
                eng.stream_partition, eng.params_op = lay.stream_partition, object()
                d.eng, d.rank, d.world, d.overlap_decoder, d.mode, d.stream = eng, rank, world, overlap, "p2p", stream
                rr = d.ranges()
                # streamed: enc1 | enc2 | ench | zh ch dec1 | dec2 dec3 decx + prior tables
                assert rr[0][0] == 0 and rr[-1][1] == lay.n_params and len(rr) == (5 if stream else 3 if overlap else 1)
                assert all(rr[i][1] == rr[i + 1][0] for i in range(len(rr) - 1))
                for i in range(len(rr)):
                    b, e = d.range_shard(i)
                    assert b % 4 == 0 and e % 4 == 0 and rr[i][0] <= b <= e <= rr[i][1]
                    covered[b:e] += 1
            assert (covered == 1).all()


def test_stream_plan_covers_every_block_once():
    """Layout.stream_plan / stream_partition (host logic of the streamed update): every block is in at most one
    segment, segments are contiguous ranges, and only the first encoder layer is left for the end of the step."""
    from dmvae_b200.engine import Layout
    for kw in (dict(model="dmvae", trunk=(500, 500), head=2000, decoder=(2000, 500, 500)),
               dict(model="vade", trunk=(2000, 500, 500), head=0, decoder=(500, 500, 2000)),
               dict(model="dmvae", trunk=(500, 500), head=2000, decoder=(2000,))):
        lay = Layout(input_dim=784, latent_dim=10, n_classes=10, name="m", **kw)
        plan = lay.stream_plan()
        names = [n for v in plan.values() for n in v]
        assert len(names) == len(set(names))
        assert set(names) == (set(lay.layers) | {"priors"}) - {lay.enc_chain[0]}
        part = lay.stream_partition()
        assert part[0][0] == 0 and part[-1][1] == lay.n_params
        assert all(part[i][1] == part[i + 1][0] for i in range(len(part) - 1))
        for v in plan.values():
            for off, n in lay.merged_ranges(v):
                assert (off, off + n) in part
        e1 = lay.layers[lay.enc_chain[0]]
        assert (e1.offset, e1.offset + e1.size) in part


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dmvae_b200 import dp
    from oracle import reference_graph as rg
    n = 1000
    rs = np.random.RandomState(0)
    theta0 = rs.randn(n)
    g_all = rs.randn(world, n)                       # per-rank partial gradients (already scaled by 1/global batch)
    results = {}
    for mode in ("allreduce", "sharded"):
        theta = torch.tensor(theta0.copy())
        m, v = np.zeros(n), np.zeros(n)
        grads = torch.tensor(g_all[rank].copy())

        def apply_adam(b, e, gsum):
            th = theta.numpy()[b:e]
            rg.adam_tf_step(th, gsum.numpy().copy(), m[b:e], v[b:e], 1, 0.002)

        if mode == "allreduce":
            dp.allreduce_adam_reference(grads, apply_adam)
        else:
            dp.sharded_adam_reference(grads, theta, apply_adam, rank, world)
        results[mode] = theta.numpy().copy()
    ref = theta0.copy()
    rg.adam_tf_step(ref, g_all.sum(0), np.zeros(n), np.zeros(n), 1, 0.002)
    ok = np.allclose(results["allreduce"], ref, atol=1e-12) and np.allclose(results["sharded"], ref, atol=1e-12)
    out.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_sharded_update_equals_allreduce_update_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
