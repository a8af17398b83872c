import sys; sys.path.insert(0, '.')
import numpy as np, torch
sys.path.insert(0, "tests")
from test_gpu_model import _make, _data, relerr, rel_l2
from oracle import reference_graph as rg
for model, L, K, B in (("dmvae", 10, 10, 256), ("vade", 64, 50, 512)):
    cfg, eng, V = _make(model, "bf16", L=L, K=K, B=B)
    X, eps, gum = _data(B, 784, L, K)
    out, g = rg.loss_and_grads(cfg, V, X, eps)
    outb, gb = rg.loss_and_grads(cfg, V, X, eps, gemm_round=rg.bf16_round)
    eng.forward_backward(torch.tensor(X, device="cuda"), B, torch.tensor(eps, device="cuda"), None, 1.0)
    torch.cuda.synchronize()
    print(model, "loss", eng.loss_out.cpu().numpy(), out["loss"], outb["loss"])
    ps = eng.per_sample[:B].cpu().numpy()
    for i, k in enumerate(["recon_ps", "kl_c_ps", "kl_z_ps", "elbo_ps"]):
        print("  %-9s max rel err vs fp64 %.3g   vs bf16-emu %.3g" % (k, np.max(np.abs(ps[:, i] - out[k]) / (np.abs(out[k]) + 1e-2)),
              np.max(np.abs(ps[:, i] - outb[k]) / (np.abs(outb[k]) + 1e-2))))
    for name in rg.trainable_names(cfg):
        got = eng.get_variable(name, grad=True)
        print("  %-50s l2 %.4f max %.4f | emu: l2 %.4f max %.4f | emu-vs-fp64 l2 %.4f" % (name, rel_l2(got, g[name]), relerr(got, g[name]),
              rel_l2(got, gb[name]), relerr(got, gb[name]), rel_l2(gb[name], g[name])))
