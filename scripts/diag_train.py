import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dmvae_b200 import base_models, nn
from dmvae_b200.session import Session
from dmvae_b200.includes.utils import Dataset
rs = np.random.RandomState(0)
protos = (rs.uniform(size=(10, 784)) < 0.2)
cls = np.arange(2000) % 10
X = (protos[cls] ^ (rs.uniform(size=(2000, 784)) < 0.03)).astype(np.float32)
for gd in ("fp32", "bf16"):
    model = base_models.DeepMixtureVAE("dmvae", "binary", 784, 10, 10, activation=nn.relu,
                                       initializer=nn.xavier_initializer).build_graph()
    model.gemm_dtype = gd
    model.define_train_step(0.002, 100)
    sess = Session()
    data = Dataset((X, cls), batch_size=100)
    losses = [model.train_op(sess, data, 1.0) for _ in range(4)]
    print(gd, os.environ.get("DMVAE_CHAIN"), os.environ.get("DMVAE_GRAPH"), losses)
