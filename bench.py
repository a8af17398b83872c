#!/usr/bin/env python
"""Benchmark of the DMVAE training step (BASELINE.json: "DMVAE train samples/sec (fwd+bwd ELBO)").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload = BASELINE.json configs[1]: DMVAE, MNIST-shaped binarised 784-d synthetic data, K=10 clusters, latent 10,
batch 4096 per GPU, bf16 tcgen05 GEMMs.  One step = encoder -> Philox reparameterisation -> decoder -> fused ELBO
fwd+bwd -> gradient GEMMs -> (gradient reduction over NVLink for N>1) -> Adam.

`value`      : samples/s with the batches already resident in HBM (CUDA events, max over ranks).
`e2e`        : same metric through the public API (engine.run_epoch, the body of model.train_op) with the data in
               pinned HOST memory: per step an H2D copy of the batch and a D2H read of the loss inside the timed region.
`roofline`   : the fused ELBO kernel (HBM-bound): algorithmic bytes per launch / mean CUDA-event duration of that
               launch inside the timed region, against MEASURED_PEAKS.json.
`cpu_baseline`: the CPU restatement of the reference step (oracle/, kind "port": TensorFlow 1.x is not installable)
               timed on this box's host cores on a bounded sample.
--impl reference times that CPU restatement on the same config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "DMVAE train samples/sec (fwd+bwd ELBO)"
UNIT = "samples/s"
D, L, K = 784, 10, 10
BATCH_PER_GPU = 4096
TRUNK, HEAD, DEC = (500, 500), 2000, (2000, 500, 500)
FLOP_PER_SAMPLE = 25.40e6                      # 6*sum(Kin*Nout) - 2*D*H1
ELBO_BYTES_PER_SAMPLE = (1 + 2 + 2) * D + 4 * (3 * L + K) + 4 * (2 * L + K) + 4 * K + 4 * L + 12   # u8 X, bf16 logits/grad: 4292


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


def synth_batches(n_rows, seed=1):
    import numpy as np
    rng = np.random.RandomState(seed)
    return (rng.uniform(size=(n_rows, D)) < 0.1307).astype(np.uint8)


def dbg(msg):
    if os.environ.get("DMVAE_BENCH_DEBUG"):
        print("[bench %s] %s" % (os.environ.get("RANK", "0"), msg), file=sys.stderr, flush=True)


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from dmvae_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node %d" % args.gpus
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = BATCH_PER_GPU
    eng = Engine(model="dmvae", input_type="binary", input_dim=D, latent_dim=L, n_classes=K, trunk=TRUNK, head=HEAD,
                 decoder=DEC, name="dmvae", gemm_dtype="bf16", max_rows=B, device=dev, seed=0)
    dp = None
    if world > 1:
        from dmvae_b200.dp import DataParallel
        dp = DataParallel(eng, mode=args.dp_mode)
    opt = eng.optimizer("train", 0.002)
    NB = 16                                                         # resident batches rotated through (51 MB of u8)
    host = torch.from_numpy(synth_batches(NB * B, seed=1 + rank)).pin_memory()
    resident = host.to(dev)
    K_, W_ = args.steps, args.warmup

    xs = torch.empty(B, D, dtype=torch.uint8, device=dev)           # static input buffer of the captured step

    def step(i):
        xs.copy_(resident[(i % NB) * B:(i % NB + 1) * B], non_blocking=True)     # device-to-device, 3.2 MB
        if dp is None:
            eng.train_step(xs, B, opt)
        else:
            dp.train_step(xs, B, opt)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(dev)

    dbg("setup done (dp mode %s)" % (dp.mode if dp else None))
    for i in range(W_):
        step(i)
    barrier()
    dbg("warmup done")
    # ---- timed region 1: inputs resident in HBM ----
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    l0 = eng.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(K_):
        step(W_ + i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    dbg("region 1 done: %.3f ms/step" % (ms / K_))
    launches = eng.launches() - l0
    # ---- timed region 2: end to end from pinned host memory through the public epoch loop ----
    NBH = min(K_, 128)                                              # host-resident batches of the e2e pass (<= 411 MB pinned)
    if NBH > NB:
        host = torch.from_numpy(synth_batches(NBH * B, seed=1 + rank)).pin_memory()
    eng.run_epoch(host, B, opt, max_steps=min(W_, NBH)) if dp is None else dp.run_epoch(host, B, opt, max_steps=min(W_, NBH))
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    done = 0
    while done < K_:
        n = min(NBH, K_ - done)
        (eng.run_epoch(host, B, opt, max_steps=n) if dp is None else dp.run_epoch(host, B, opt, max_steps=n))
        done += n
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    dbg("region 2 done: %.3f ms/step" % (ms_e2e / K_))
    clk = clocks.stop() if rank == 0 else None
    # ---- timed region 3: per-kernel device times.  Each kernel is launched `n_inst` times back to back inside one
    #      CUDA graph on this rank's own step buffers and the replay is bracketed by CUDA events on the launching
    #      stream (an event pair around a single eager launch would mostly measure the host's launch gap).  The
    #      ELBO operands are L2-warm, as they are in the step (the decoder GEMM has just written them). ----
    import ctypes as C
    from dmvae_b200 import _abi
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    from gemm_bench import time_gemms
    n_inst = 20

    def graph_time_us(fn):
        fn()
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(n_inst):
                fn()
        g.replay()
        torch.cuda.synchronize(dev)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        g.replay()
        t1.record()
        torch.cuda.synchronize(dev)
        return t0.elapsed_time(t1) * 1e3 / n_inst

    ea = eng._elbo_args(xs, _abi.U8, B, 1.0, 1.0 / (world * B))
    elbo_us = graph_time_us(lambda: _abi.check(eng.lib.dmvae_elbo_fwd_bwd(eng.ctx, C.byref(ea), eng._stream())))
    adam_us = None
    if dp is None:
        adam_us = graph_time_us(lambda: _abi.check(eng.lib.dmvae_adam(
            eng.ctx, eng.params.data_ptr(), eng.grads.data_ptr(), opt.m.data_ptr(), opt.v.data_ptr(),
            eng.params_op.data_ptr(), eng.n_params, 1e-9, None, opt.beta1, opt.beta2, opt.eps, 1.0, 1, eng._stream())))
    # the same kernel at 16 batches' worth of rows (65 536): the launch + single-wave cost that dominates at 4096 rows
    # amortises, which is the figure to read against the HBM roofline
    elbo_big_rows, elbo_big_us = 16 * B, None
    if world == 1:
        from elbo_bench import time_elbo
        elbo_big_us, _ = time_elbo(eng.lib, eng.ctx, elbo_big_rows, D, L, K, n_inst)
    gemm_us_step, _, _ = time_gemms(eng.lib, eng.ctx, B, n_inst, verbose=False, dev=dev)
    gemm_ms_step = gemm_us_step * 1e-3
    dbg("region 3 done")

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        _finish(world, dev)
        return
    pk = peaks()
    value = world * B * K_ / (ms * 1e-3)
    e2e = world * B * K_ / (ms_e2e * 1e-3)
    elbo_avg_ms = elbo_us * 1e-3
    elbo_bytes = ELBO_BYTES_PER_SAMPLE * B
    achieved = elbo_bytes / (elbo_avg_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "elbo_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K_, "warmup": W_,
        "ms_per_step": ms / K_, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": "DMVAE MNIST-shaped binarised 784-d, K=10, latent 10, batch 4096 per GPU, bf16 GEMMs "
                               "(BASELINE.json configs[1])",
                   "batch_per_gpu": B, "global_batch": world * B, "hidden": "784-500-500-2000 / 2000-500-500-784",
                   "parallelism": "dp%d" % world if world > 1 else "single",
                   "exchange": None if dp is None else (dp.mode + ("+nvls" if getattr(dp, "_mc", 0) else "")),
                   "l2": "per-step working set ~240 MB (> 126 MB L2); inputs rotate over 16 resident batches",
                   "noise": "device Philox4x32-10", "optimizer": "Adam (TF semantics), every step"},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": world * B * D, "d2h_bytes_per_step": world * 16,
                "ms_per_step": ms_e2e / K_, "api": "Engine.run_epoch (body of model.train_op) from pinned host uint8"},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": {"kernel": "elbo_rowtile_kernel<u8,bf16,binary> (fused ELBO fwd+bwd)", "bound": "hbm", "achieved": achieved,
                     "peak": pk["hbm"], "unit": "GB/s", "frac": achieved / pk["hbm"], "traffic": traffic,
                     "peak_source": pk["src"], "bytes_per_launch": elbo_bytes, "us_per_launch": elbo_avg_ms * 1e3,
                     "timing": "CUDA events around a graph replay of 20 launches on the step's own buffers (L2-warm, as in the step)",
                     "large_batch": None if elbo_big_us is None else {
                         "rows": elbo_big_rows, "us_per_launch": elbo_big_us,
                         "achieved": ELBO_BYTES_PER_SAMPLE * elbo_big_rows / elbo_big_us * 1e-3,
                         "frac": ELBO_BYTES_PER_SAMPLE * elbo_big_rows / elbo_big_us * 1e-3 / pk["hbm"],
                         "note": "same kernel, 65 536 rows, inputs rotating over > L2"}},
        "roofline_gemm": {"bound": "tensor", "achieved": FLOP_PER_SAMPLE * B / (gemm_ms_step * 1e-3) / 1e12,
                          "peak": pk["tf_sust"], "unit": "TFLOP/s",
                          "frac": FLOP_PER_SAMPLE * B / (gemm_ms_step * 1e-3) / 1e12 / pk["tf_sust"],
                          "gemm_ms_per_step": gemm_ms_step,
                          "note": "algorithmic FLOPs (25.40 MFLOP/sample) / sum of the device times of the step's 27 GEMM "
                                  "launches, each timed as a CUDA-graph replay of 20 back-to-back launches"},
        "adam_us": adam_us,
    }
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(bounded_seconds=20.0)
    print(json.dumps(out), flush=True)
    _finish(world, dev)


def _finish(world, dev):
    """Multi-rank teardown: captured graphs hold NCCL / symmetric-memory work, and destroying the process group under
    them can block; every rank synchronises and leaves without running the destructors."""
    if world > 1:
        import torch
        torch.cuda.synchronize(dev)
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def cpu_baseline(bounded_seconds=20.0, batch=256):
    """CPU restatement of the reference step (oracle/cpu_train.py) at the reference's CPU-runnable config
    (configs[0]: batch 256) on a bounded sample."""
    from oracle import cpu_train
    from oracle import reference_graph as rg
    cores = os.cpu_count() or 1
    cfg = rg.GraphConfig(input_dim=D, latent_dim=L, n_classes=K)
    probe = cpu_train.time_training(cfg, batch, 3, 2, cores)
    n = int(max(5, min(200, bounded_seconds / (probe["ms_per_step"] * 1e-3))))
    r = cpu_train.time_training(cfg, batch, n, 3, cores)
    return {"value": r["samples_per_s"], "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d steps of batch %d (median step %.1f ms, p10 %.1f, p90 %.1f), op-for-op fp32 PyTorch-CPU "
                      "restatement of the reference graph incl. Python batching and host noise" %
                      (r["steps"], batch, r["ms_per_step"], r["p10_ms"], r["p90_ms"])}


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path cannot run (TensorFlow 1.x is not
    installable here), so this times the op-for-op CPU port in oracle/ on the same config (batch 4096 per step),
    rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import cpu_train
    from oracle import reference_graph as rg
    cores = os.cpu_count() or 1
    cfg = rg.GraphConfig(input_dim=D, latent_dim=L, n_classes=K)
    batch = BATCH_PER_GPU
    probe = cpu_train.time_training(cfg, batch, 1, 1, cores)
    if probe["ms_per_step"] * (args.steps + args.warmup) > 240e3:      # keep the arm within a few minutes
        batch = 1024
    r = cpu_train.time_training(cfg, batch, args.steps, args.warmup, cores)
    val = r["mean_samples_per_s"]
    sample = "%d steps x %d samples on %d host threads (PyTorch-CPU fp32 port of the TF graph)" % (r["steps"], batch, cores)
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * batch / val, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "DMVAE MNIST-shaped binarised 784-d, K=10, latent 10, batch 4096 per GPU "
                                  "(BASELINE.json configs[1]) - CPU port", "sample_batch": batch},
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dp_mode", default="auto")
    ap.add_argument("--no_cpu_baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
