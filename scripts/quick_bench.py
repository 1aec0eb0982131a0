"""Development timing sweep over tile / register-qubit choices (not the contract bench)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "qhbm-library_b200"), os.path.join(ROOT, "tests")):
  sys.path.insert(0, p)
import numpy as np
import torch
from _workloads import hea_tables
from qhbmlib import engine


def run(n, layers, u, T, K, grad, ham="xxz", reps=3):
  gates, nsym, terms, offs = hea_tables(n, layers, ham)
  plan = engine.ExpectationPlan(gates, n, nsym, terms, offs, grad, T, K)
  rng = np.random.default_rng(0)
  phi = torch.tensor(rng.uniform(-1, 1, nsym).astype(np.float32), device="cuda")
  basis = torch.tensor(rng.choice(1 << n, u, replace=False).astype(np.int64), device="cuda")
  dg = torch.tensor(rng.uniform(0, 1, (u, 1)).astype(np.float32), device="cuda")
  f = (lambda: plan.forward_adjoint(basis, phi, dg)) if grad else (lambda: plan.forward(basis, phi))
  f()
  torch.cuda.synchronize()
  best = 1e9
  for _ in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    f()
    b.record()
    torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b))
  print(f"n={n} L={layers} U={u} T={plan.info['tile_qubits']} K={plan.info['reg_qubits']} grad={grad} "
        f"chunk={plan.info['chunk']} launches={plan.info['launches']} passes={plan.info['passes']}: "
        f"{best:.2f} ms  {u / best * 1e3:.0f} bitstrings/s", flush=True)


if __name__ == "__main__":
  for T, K in [(13, 4), (12, 4), (13, 5), (12, 5)]:
    run(16, 2, 4096, T, K, True)
  for T, K in [(13, 5), (14, 5), (13, 4)]:
    run(16, 2, 4096, T, K, False)
  run(12, 2, 4096, 0, 4, True, "tfim")
  run(12, 2, 4096, 0, 5, True, "tfim")
  run(20, 2, 256, 13, 5, False, "tfim")
  run(20, 2, 256, 14, 5, False, "tfim")
