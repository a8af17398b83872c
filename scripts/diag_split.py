import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dmvae_b200.engine import Engine
def make(rows):
    return Engine(model="dmvae", input_type="binary", input_dim=784, latent_dim=10, n_classes=10, trunk=(500, 500), head=2000,
                  decoder=(2000, 500, 500), name="dmvae", gemm_dtype="fp32", max_rows=rows, seed=0)
rs = np.random.RandomState(1)
X = torch.tensor((rs.uniform(size=(256, 784)) < 0.1307).astype(np.uint8), device="cuda")
full = make(256); full.overlap = False
full.forward_backward(X, 256, inv_global_batch=1/256, row_offset=0)
torch.cuda.synchronize()
names = ["dmvae/encoder_network/dense/kernel", "dmvae/encoder_network/dense_1/kernel", "dmvae/decoder_network/dense/kernel"]
gf = {n: full.get_variable(n, grad=True).astype(np.float64) for n in names}
half = make(128); half.overlap = False
gs = {n: 0 for n in names}
for r in range(2):
    half.forward_backward(X[r*128:(r+1)*128], 128, inv_global_batch=1/256, row_offset=r*128)
    torch.cuda.synchronize()
    for n in names: gs[n] = gs[n] + half.get_variable(n, grad=True).astype(np.float64)
for n in names:
    d = np.abs(gs[n] - gf[n]); i = np.unravel_index(np.argmax(d), d.shape)
    print(n, "max abs diff %.3e at %s: full %.6e split %.6e; max|g| %.3e; n(|d|>1e-7*max)=%d" % (d.max(), i, gf[n][i], gs[n][i], np.abs(gf[n]).max(), (d > 1e-6*np.abs(gf[n]).max()).sum()))
print("g[396,379] full %.6e split %.6e" % (gf[names[0]][396,379], gs[names[0]][396,379]))
