"""Times the per-state-symbol entry points against the shared-row ones on config 3's shape
(16 qubits, HEA L=2, XXZ ring, 4096 states, forward + adjoint): same sweeps, U coefficient tables instead of one."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "qhbm-library_b200"), os.path.join(ROOT, "tests")):
  sys.path.insert(0, p)
from _workloads import hea_tables  # noqa: E402
from qhbmlib import engine  # noqa: E402

n, u = 16, 4096
gates, nsym, terms, offs = hea_tables(n, 2, "xxz")
plan = engine.ExpectationPlan(gates, n, nsym, terms, offs, True)
rng = np.random.default_rng(3)
phi = torch.tensor(rng.uniform(-1, 1, nsym).astype(np.float32), device="cuda")
rows = phi[None, :].repeat(u, 1).contiguous()
basis = torch.tensor(rng.choice(1 << n, u, replace=False).astype(np.int64), device="cuda")
dg = torch.full((u, 1), 1.0 / u, device="cuda")


def timed(sym, reps=5):
  for _ in range(3):
    plan.forward_adjoint(basis, sym, dg)
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(reps):
    e, g = plan.forward_adjoint(basis, sym, dg)
  b.record()
  torch.cuda.synchronize()
  return a.elapsed_time(b) / reps, e, g


t0, e0, g0 = timed(phi)
t1, e1, g1 = timed(rows)
print(json.dumps({"workload": "config 3 shape, 4096 states, forward + adjoint", "ms_shared_row": t0, "ms_symbol_rows": t1,
                  "max_abs_diff_expectation": float((e0 - e1).abs().max()),
                  "max_abs_diff_gradient": float((g0 - g1).abs().max())}))
