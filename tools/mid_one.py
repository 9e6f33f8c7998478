"""One likelihood and one prediction launch of the generic kernel (N points, B objects) -- the command profiled by ncu
for the 65..224-point path.  python tools/mid_one.py [N] [B]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cosmogp_b200.batch import DeviceBatch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 224
b = int(sys.argv[2]) if len(sys.argv) > 2 else 8000
rng = np.random.default_rng(0)
x = np.sort(rng.uniform(0, n / 2.0, (b, n)), axis=1); y = rng.standard_normal((b, n)); ye = np.full((b, n), 0.2)
bt = DeviceBatch(x.ravel(), y.ravel(), np.arange(b + 1) * n, y_err=ye.ravel())
g = torch.from_numpy(np.linspace(0, n / 2.0, 100)).cuda()
ll, info = bt.ll_dev([0.7, 2.0], 0.0)
m, v, _ = bt.predict_dev([0.7, 2.0], 0.0, g, None, None, True)
torch.cuda.synchronize()
print(float(ll.sum()), float(m.sum()), float(v.sum()))
