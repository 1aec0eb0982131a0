"""Generates tests/golden/qhbm_golden.npz from the NumPy oracle (oracle/qhbm_oracle.py).

The reference itself cannot be imported in this environment (TensorFlow / TFQ / cirq are not
installable), so the committed vectors are oracle outputs: they freeze the oracle against drift and
give the GPU tests fixed numbers to hit.  Re-run with:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import qhbm_oracle as orc  # noqa: E402

CASES = [  # (name, n, layers, hamiltonian, n_states, seed)
    ("c1_tfim4", 4, 2, "tfim", 16, 5),
    ("tfim8", 8, 2, "tfim", 6, 8),
    ("xxz10", 10, 3, "xxz", 4, 10),
    ("c2_tfim12", 12, 2, "tfim", 4, 12),
    ("c3_xxz16", 16, 2, "xxz", 2, 16),
]


def main():
  out = {}
  for name, n, layers, ham, n_states, seed in CASES:
    rng = np.random.default_rng(seed)
    gates, names = orc.hea_circuit(n, layers)
    phi = rng.uniform(-1, 1, len(names)).astype(np.float32)
    ops = [orc.tfim_ring(n) if ham == "tfim" else orc.xxz_ring(n)] + orc.kobe_shards(n, 2)[:3]
    basis = np.arange(1 << n) if n_states == 1 << n else rng.choice(1 << n, n_states, replace=False)
    dg = rng.uniform(-1, 1, (len(basis), len(ops))).astype(np.float32)
    for mode in ("exact", "tfq_fd"):
      e, g = orc.batch_expectation_and_gradient(gates, n, phi, basis, ops, dg, mode)
      out[f"{name}/{mode}/grad"] = g
    out[f"{name}/phi"], out[f"{name}/basis"], out[f"{name}/dgrad"], out[f"{name}/exp"] = phi, basis, dg, e
    print(name, "done")
  # EBM side
  rng = np.random.default_rng(99)
  th = rng.normal(0, 0.3, len(orc.parity_indices(10, 2))).astype(np.float32)
  en = orc.kobe_energy(orc.all_bitstrings(10), 2, th)
  out["kobe10/theta"], out["kobe10/energies"] = th, en
  out["kobe10/logz_entropy"] = np.array([orc.analytic_log_partition(en), orc.analytic_entropy(en)])
  np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "qhbm_golden.npz"), **out)


if __name__ == "__main__":
  main()
