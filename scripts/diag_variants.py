import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dmvae_b200.engine import Engine
def make(rows, dt="fp32"):
    return Engine(model="dmvae", input_type="binary", input_dim=784, latent_dim=10, n_classes=10, trunk=(500, 500), head=2000,
                  decoder=(2000, 500, 500), name="dmvae", gemm_dtype=dt, max_rows=rows, seed=0)
rs = np.random.RandomState(1)
Xg = (rs.uniform(size=(3, 256, 784)) < 0.1307).astype(np.uint8)
res = {}
for overlap in (False, True):
    for graphs in (False, True):
        for static in (False, True):
            e = make(256); e.overlap = overlap; e.use_graphs = graphs
            o = e.optimizer("train", 0.002)
            xs = torch.empty(256, 784, dtype=torch.uint8, device="cuda")
            for i in range(3):
                if static:
                    xs.copy_(torch.tensor(Xg[i], device="cuda")); e.train_step(xs, 256, o)
                else:
                    e.train_step(torch.tensor(Xg[i], device="cuda"), 256, o)
            torch.cuda.synchronize()
            res[(overlap, graphs, static)] = e.get_variable("dmvae/encoder_network/dense/kernel")
            e.close()
base = res[(False, False, True)]
for k, v in res.items():
    print("overlap=%s graphs=%s static=%s: max diff vs (no overlap, eager, static) %.3e" % (k + (np.abs(v - base).max(),)))
