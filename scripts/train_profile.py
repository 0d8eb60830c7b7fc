"""Two train steps of the config-4 shape (64 patches of 48x48, x4) for an `ncu --metrics gpu__time_duration.sum` launch
list; the second step is the warm one."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dcscn-super-resolution_b200"))
from helper import engine as E, tf_bundle
r = tf_bundle.BundleReader(os.path.join(ROOT, "tests", "golden", "models", "dcscn_L12_F196to48_Sc4_NIN_A64_PS_R1F32.ckpt"))
eng = E.Engine(E.make_config(scale=4, dropout_keep=0.8)); eng.set_params({k: r.get_tensor(k) for k in r.keys()})
g = torch.Generator().manual_seed(2)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
x = (torch.rand(n, 48, 48, 1, generator=g) * 255).cuda(); x2 = (torch.rand(n, 192, 192, 1, generator=g) * 255).cuda(); y = (torch.rand(n, 192, 192, 1, generator=g) * 255).cuda()
for i in range(3):
    print(eng.train_step(x, x2, y, 0.002, i, apply_update=True))
torch.cuda.synchronize()
