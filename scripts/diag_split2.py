import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dmvae_b200.engine import Engine
def make(rows):
    e = Engine(model="dmvae", input_type="binary", input_dim=784, latent_dim=10, n_classes=10, trunk=(500, 500), head=2000,
                  decoder=(2000, 500, 500), name="dmvae", gemm_dtype="fp32", max_rows=rows, seed=0)
    e.use_graphs = False; e.overlap = False
    return e
rs = np.random.RandomState(1)
Xg = (rs.uniform(size=(3, 256, 784)) < 0.1307).astype(np.uint8)
full, half = make(256), make(128)
of, oh = full.optimizer("t", 0.002), half.optimizer("t", 0.002)
nm = "dmvae/encoder_network/dense/kernel"
for i in range(3):
    X = torch.tensor(Xg[i], device="cuda")
    full.forward_backward(X, 256, inv_global_batch=1/256, row_offset=0)
    gfull = full.grads.clone()
    acc = torch.zeros_like(half.grads)
    zs = []
    for r in range(2):
        half.step_count = full.step_count
        half.forward_backward(X[r*128:(r+1)*128], 128, inv_global_batch=1/256, row_offset=r*128)
        acc += half.grads
        zs.append(half.zh[:128].clone())
    torch.cuda.synchronize()
    d = (acc - gfull).abs()
    print("step %d: grad max abs diff %.3e (max|g| %.3e); zh diff %.3e; loss full %.6f" % (i, float(d.max()), float(gfull.abs().max()),
          float((torch.cat(zs) - full.zh[:256]).abs().max()), float(full.loss_out[3])))
    gk_f = full.get_variable(nm, grad=True); half.grads.copy_(acc); gk_h = half.get_variable(nm, grad=True)
    dd = np.abs(gk_f - gk_h); print("   enc1 kernel grad: max diff %.3e, n(>1e-8)=%d" % (dd.max(), (dd > 1e-8).sum()))
    full.adam(of); half.adam(oh); full.step_count += 1
    torch.cuda.synchronize()
    dw = np.abs(full.get_variable(nm) - half.get_variable(nm))
    print("   after adam: enc1 kernel max diff %.3e n(>5e-6)=%d ; flat params max diff %.3e" % (dw.max(), (dw > 5e-6).sum(), float((full.params - half.params).abs().max())))
