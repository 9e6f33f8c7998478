"""Tiny run through every kernel family (for compute-sanitizer memcheck / racecheck)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cosmogp_b200 as cg
from cosmogp_b200.batch import DeviceBatch
from cosmogp_b200 import dense
rng = np.random.default_rng(0)
for n, b in ((60, 40), (37, 9), (100, 3)):
    x = np.sort(rng.uniform(0, 30, (b, n)), axis=1); y = rng.standard_normal((b, n)); ye = rng.uniform(0.1, 0.3, (b, n))
    bt = DeviceBatch(x.ravel(), y.ravel(), np.arange(b + 1) * n, y_err=ye.ravel())
    print(n, bt.log_likelihood([0.7, 2.0], 0.05)[0])
    grid = np.linspace(0, 30, 45)
    m, v, _ = bt.predict([0.7, 2.0], 0.05, grid); print(m.sum(), v.sum())
    print(bt.loo([0.7, 2.0], 0.05, mode=1, mean=np.zeros(b * n))[2].sum())
    if n <= 64:
        fac = bt.factor_dev([0.7, 2.0], 0.05)
        mm, vv, _ = bt.predict_factored_dev(fac, torch.from_numpy(grid).cuda(), None, None, True); print(float(mm.sum()))
        print(bt.ll_objhyp(np.tile([0.7, 2.0], (b, 1)), np.arange(b))[0].sum())
    k, ki, _ = bt.matrices([0.7, 2.0], 0.05); print(k[0].sum(), ki[0].sum())
x2 = rng.uniform(-50, 50, (300, 2)); y2 = rng.standard_normal(300)
obj = dense.LargeObject(x2, y2, np.full(300, 0.2), None, dim=2)
print(obj.factor([1.0, 30.0, 25.0, 50.0], 0.0)); print(obj.predict(rng.uniform(-50, 50, (200, 2)))[1].sum())
print(cg.cholesky_inverse(np.eye(5) * 2.0)[0, 0])
print("sanitize run ok")
