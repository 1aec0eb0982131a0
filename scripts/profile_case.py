"""One invocation of the hot path for ncu: python scripts/profile_case.py n layers U T K grad [ham]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "qhbm-library_b200"), os.path.join(ROOT, "tests")):
  sys.path.insert(0, p)
import numpy as np
import torch
from _workloads import hea_tables
from qhbmlib import engine

n, layers, u, T, K, grad = (int(x) for x in sys.argv[1:7])
ham = sys.argv[7] if len(sys.argv) > 7 else "xxz"
reps = int(sys.argv[8]) if len(sys.argv) > 8 else 1
gates, nsym, terms, offs = hea_tables(n, layers, ham)
plan = engine.ExpectationPlan(gates, n, nsym, terms, offs, bool(grad), T, K)
rng = np.random.default_rng(0)
phi = torch.tensor(rng.uniform(-1, 1, nsym).astype(np.float32), device="cuda")
basis = torch.tensor(rng.choice(1 << n, u, replace=False).astype(np.int64), device="cuda")
dg = torch.tensor(rng.uniform(0, 1, (u, len(offs) - 1)).astype(np.float32), device="cuda")
best = None
for _ in range(reps):
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  if grad:
    plan.forward_adjoint(basis, phi, dg)
  else:
    plan.forward(basis, phi)
  b.record()
  torch.cuda.synchronize()
  t = a.elapsed_time(b)
  best = t if best is None else min(best, t)
print(plan.info, f"best {best:.2f} ms {u / best * 1e3:.0f} bitstrings/s")
