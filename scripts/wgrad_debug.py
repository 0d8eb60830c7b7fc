import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dcscn-super-resolution_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from helper import engine as E
import dcscn_oracle as O
kw = dict(scale=2, layers=3, filters=24, min_filters=16, filters_decay_gamma=1.5, nin_filters=16, nin_filters2=16)
cfg = O.OracleConfig(**kw)
wts = {k: v.astype(np.float32) for k, v in O.he_init_weights(cfg, seed=0).items()}
g = np.random.RandomState(1)
n, h, w = 2, 12, 10
x = (g.rand(n, h, w, 1) * 255).astype(np.float32); x2 = (g.rand(n, 2*h, 2*w, 1) * 255).astype(np.float32)
y = np.clip(x2 + g.randn(n, 2*h, 2*w, 1) * 10, 0, 255).astype(np.float32)
res = []
for impl in (1, 0):
    eng = E.Engine(E.make_config(dropout_keep=1.0, **kw)); eng.set_params(wts); eng.set_option("wgrad_impl", impl)
    print("impl", impl, eng.train_step_host(x, x2, y, lr=0.002, seed=1, apply_update=False))
    res.append({k: eng.get_grad(k) for k in wts}); eng.close()
for k in wts:
    if k.endswith("conv_W"): print(k, np.abs(res[0][k]-res[1][k]).max(), np.abs(res[0][k]).max())
