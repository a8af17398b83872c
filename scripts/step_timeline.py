#!/usr/bin/env python
"""Kernel timeline of the captured training step (torch.profiler / CUPTI): per-kernel start, duration and stream for
one CUDA-graph replay, so that gaps and overlap between the dgrad chain and the side stream can be read off."""
import json
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from dmvae_b200.engine import Engine  # noqa: E402


def main():
    """Single GPU: python scripts/step_timeline.py.  Data parallel: torchrun --nproc-per-node N scripts/step_timeline.py
    (rank 0 prints its own timeline, which then includes the cross-GPU barriers and the exchange kernel)."""
    B = bench.BATCH_PER_GPU
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    eng = Engine(model="dmvae", input_type="binary", input_dim=bench.D, latent_dim=bench.L, n_classes=bench.K, trunk=bench.TRUNK,
                 head=bench.HEAD, decoder=bench.DEC, name="dmvae", gemm_dtype="bf16", max_rows=B, seed=0)
    if world > 1:
        from dmvae_b200.dp import DataParallel
        DataParallel(eng, mode="auto")
    opt = eng.optimizer("train", 0.002)
    xs = torch.from_numpy(bench.synth_batches(B, seed=1 + rank)).cuda()
    for _ in range(5):
        eng.train_step(xs, B, opt)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            eng.train_step(xs, B, opt)
        torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
        if rank != 0:
            torch.cuda.synchronize()
            os._exit(0)
    out = os.path.join(ROOT, "gpurun_out", "step_trace.json")
    prof.export_chrome_trace(out)
    ev = [e for e in json.load(open(out))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    ev.sort(key=lambda e: e["ts"])
    if not ev:
        print("no kernel records")
        return
    # last replay only
    step_starts = [e["ts"] for e in ev if "step_tick" in e["name"]]
    t0 = step_starts[-1]
    rows = [e for e in ev if e["ts"] >= t0]
    end = max(e["ts"] + e["dur"] for e in rows)
    print("step span %.1f us, %d kernels, sum of durations %.1f us" % (end - t0, len(rows), sum(e["dur"] for e in rows)))
    for e in rows:
        nm = e["name"].replace("(anonymous namespace)::", "").split("(")[0][:60]
        print("%8.1f %7.1f  s%-3s %s  grid=%s" % (e["ts"] - t0, e["dur"], e["args"].get("stream", "?"), nm, e["args"].get("grid", "")))
    if world > 1:
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
