"""The five BASELINE.json configurations as engine / model constructor arguments (shared by bench.py and the probes).

cfg 1: DMVAE 784-d, K=10, L=10, batch 256                      (the reference's CPU-runnable case)
cfg 2: same, batch 4096 per GPU, bf16 GEMMs                     (the headline; bench.py's default)
cfg 3: VaDE, K=50, L=64, batch 8192 GLOBAL (strong scaling at 2/4/8 GPUs; 8192 on one GPU)
cfg 4: LR mixture of experts (runLR_MOE.sh: `dmoe --classification --n_experts 16`), 784-d, batch 4096
cfg 5: DMVAE 3072-d, hidden 2000-2000-4000, K=100, L=128, 8192 per GPU (65 536 global on 8 GPUs)
"""
from __future__ import annotations

CONFIGS = {
    1: dict(name="cfg1", model="dmvae", D=784, L=10, K=10, trunk=(500, 500), head=2000, decoder=(2000, 500, 500),
            batch=256, scaling="weak", binarised=True,
            workload="DMVAE default train.py shapes, synthetic MNIST-shaped binarised 784-d, K=10, latent 10, batch 256 "
                     "(BASELINE.json configs[0])"),
    2: dict(name="cfg2", model="dmvae", D=784, L=10, K=10, trunk=(500, 500), head=2000, decoder=(2000, 500, 500),
            batch=4096, scaling="weak", binarised=True,
            workload="DMVAE MNIST-shaped binarised 784-d, K=10, latent 10, batch 4096 per GPU, bf16 GEMMs "
                     "(BASELINE.json configs[1])"),
    3: dict(name="cfg3", model="vade", D=784, L=64, K=50, trunk=(2000, 500, 500), head=0, decoder=(500, 500, 2000),
            batch=8192, scaling="strong", binarised=True,
            workload="VaDE GMM prior, MNIST-shaped binarised 784-d, K=50, latent 64, batch 8192 global "
                     "(BASELINE.json configs[2])"),
    4: dict(name="cfg4", model="dmoe", D=784, L=1, K=16, trunk=(500, 500), head=2000, decoder=(2000, 500, 500),
            batch=4096, scaling="weak", binarised=True, n_experts=16, output_dim=10,
            workload="LR mixture of experts (runLR_MOE.sh: dmoe --classification), 16 logistic-regression experts gated by "
                     "the discrete VAE's q(c|x), synthetic 784-d, batch 4096 (BASELINE.json configs[3])"),
    5: dict(name="cfg5", model="dmvae", D=3072, L=128, K=100, trunk=(2000, 2000), head=4000, decoder=(4000, 2000, 2000),
            batch=8192, scaling="weak", binarised=False,
            workload="DMVAE CIFAR-shaped 3072-d (soft targets), hidden 2000-2000-4000, K=100, latent 128, batch 8192 per GPU "
                     "= 65 536 global on 8 GPUs (BASELINE.json configs[4])"),
}


def gemm_flop_per_sample(cfg) -> float:
    """fwd+bwd GEMM FLOPs per sample = 6 * sum(K_in * N_out) - 2 * D * H1 (no dgrad of the first layer)."""
    D, L, K = cfg["D"], cfg["L"], cfg["K"]
    dims = []
    if cfg["model"] in ("dmvae", "dmoe"):
        h1, h2 = cfg["trunk"]
        hh = cfg["head"]
        dims += [(D, h1), (h1, h2), (h2, hh), (h2, hh), (hh, L), (hh, L), (hh, K)]
        first = D * h1
    else:
        prev = D
        for h in cfg["trunk"]:
            dims.append((prev, h))
            prev = h
        dims += [(prev, L), (prev, L)]
        first = D * cfg["trunk"][0]
    if cfg["model"] == "dmoe":
        # DeepMoE (models.py:240-250): lossVAE=0 - only trunk, c-head and the experts are trained / evaluated
        h1, h2 = cfg["trunk"]
        hh = cfg["head"]
        dims = [(D, h1), (h1, h2), (h2, hh), (hh, K), (D, cfg["n_experts"] * cfg["output_dim"])]
        return 6.0 * sum(a * b for a, b in dims) - 2.0 * first - 2.0 * D * cfg["n_experts"] * cfg["output_dim"]
    prev = L
    for h in cfg["decoder"]:
        dims.append((prev, h))
        prev = h
    dims.append((prev, D))
    return 6.0 * sum(a * b for a, b in dims) - 2.0 * first


def elbo_bytes_per_sample(cfg, x_bytes: int, logit_bytes: int) -> int:
    """read X + read decoder logits + write d_decoded, plus the latent I/O
    4(3L+K) + 4(2L+K) + 4K + 4L + 12."""
    D, L, K = cfg["D"], cfg["L"], cfg["K"]
    return (x_bytes + 2 * logit_bytes) * D + 4 * (3 * L + K) + 4 * (2 * L + K) + 4 * K + 4 * L + 12


def make_engine(cfg, rows, device=None, gemm_dtype="bf16", seed=0):
    from dmvae_b200.engine import Engine
    model = "vade" if cfg["model"] == "vade" else "dmvae"
    moe = None
    if cfg["model"] == "dmoe":
        moe = dict(n_experts=cfg["n_experts"], output_dim=cfg["output_dim"], featLearn=False, lossVAE=False,
                   classification=True, scope="moe/moe/moe")
    eng = Engine(model=model, input_type="binary", input_dim=cfg["D"], latent_dim=cfg["L"], n_classes=cfg["K"],
                 trunk=cfg["trunk"], head=cfg["head"], decoder=cfg["decoder"], name=model, gemm_dtype=gemm_dtype,
                 max_rows=rows, device=device, seed=seed, moe=moe)
    eng.x_scale = x_scale(cfg)
    return eng


def synth_inputs(cfg, n_rows, seed=1):
    """SURVEY 8(d) synthetic inputs in STORAGE form (uint8): binarised 1{u < 0.1307} as 0/1, or 8-bit intensities
    randint(0,256) whose value is byte * x_scale(cfg) = byte / 255 (soft targets, includes/utils.py:204-210)."""
    import numpy as np
    rng = np.random.RandomState(seed)
    if cfg["binarised"]:
        return (rng.uniform(size=(n_rows, cfg["D"])) < 0.1307).astype(np.uint8)
    return rng.randint(0, 256, size=(n_rows, cfg["D"])).astype(np.uint8)


def x_scale(cfg) -> float:
    return 1.0 if cfg["binarised"] else 1.0 / 255.0
