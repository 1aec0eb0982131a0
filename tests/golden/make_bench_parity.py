"""Generates tests/golden/bench_parity_<config>.npz: oracle values (numpy complex128 / float64,
oracle/qhbm_oracle.py) for bench.py's OWN synthetic inputs, so that every bench line can carry a `parity`
block without importing the oracle at run time.

  python tests/golden/make_bench_parity.py c3 [c1 c2 c3l7 c3q c4 c5]

State-vector configs: 64 bitstrings (8 at the head of each eighth of the global list, see
bench.parity_positions) -> expectations f64[64, O] and the per-state gradient of sum_j w_j <H_j> in both
gradient modes (`exact`, `tfq_fd`).  c5: log Z, entropy, 4096 logits and 64-bin probabilities of the 2^24-row
MLP sweep in float64.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "qhbm-library_b200")):
  if p not in sys.path:
    sys.path.insert(0, p)
import bench  # noqa: E402
from oracle import qhbm_oracle as orc  # noqa: E402


def state_vector(name):
  cfg = bench.CONFIGS[name]
  wl = bench.synth_workload(cfg)
  n = cfg["n"]
  pos = bench.parity_positions(len(wl["basis"]))
  basis = wl["basis"][pos]
  gates = wl["gates"].astype(orc.GATE_DTYPE)
  ops = bench.oracle_ops(wl["terms"], wl["offs"], n)
  w = wl["op_weights"].astype(np.float64)
  out = dict(positions=pos, basis=basis, op_weights=wl["op_weights"], phi=wl["phi"])
  t0 = time.time()
  if cfg["grad"]:
    e = np.zeros((len(pos), len(ops)))
    for mode in ("exact", "tfq_fd"):
      g = np.zeros((len(pos), len(wl["names"])))
      for k, idx in enumerate(basis):
        e[k], g[k] = orc.adjoint_gradient(gates, n, wl["phi"], idx, ops, w, mode)
        if k % 8 == 0:
          print(name, mode, k, f"{time.time() - t0:.0f}s", flush=True)
      out[f"grad_{mode}"] = g
    out["expectations"] = e
  else:
    out["expectations"] = orc.expectations(gates, n, wl["phi"], basis, ops)
  np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"bench_parity_{name}.npz"), **out)
  print(name, "done", f"{time.time() - t0:.0f}s")


def ebm(name):
  cfg = bench.CONFIGS[name]
  n = cfg["n"]
  rows = 1 << n
  widths, ws = bench.mlp_weights(n)
  ws = [w.astype(np.float64) for w in ws]
  rng = np.random.default_rng(99)
  logit_rows = np.unique(np.concatenate([rng.integers(0, rows, 4000), np.arange(48), rows - 1 - np.arange(48),
                                         (rows // 8) * np.arange(1, 8), (rows // 8) * np.arange(1, 8) - 1]))
  logits_all = np.empty(rows)
  for lo in range(0, rows, 1 << 18):
    idx = np.arange(lo, lo + (1 << 18), dtype=np.int64)
    bits = ((idx[:, None] >> (n - 1 - np.arange(n))[None, :]) & 1).astype(np.float64)
    h = np.tanh(np.tanh(bits @ ws[0]) @ ws[1])
    logits_all[lo:lo + (1 << 18)] = -(h @ ws[2])[:, 0]  # logits = -E (ebm.py:467-469)
  m = logits_all.max()
  e = np.exp(logits_all - m)
  s = e.sum()
  log_z = m + np.log(s)
  p = e / s
  entropy = float(-(p * (logits_all - log_z)).sum())
  np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"bench_parity_{name}.npz"),
                      log_z=log_z, entropy=entropy, logit_rows=logit_rows, logits=logits_all[logit_rows],
                      bin_probabilities=p.reshape(64, -1).sum(1))
  print(name, "log_z", log_z, "entropy", entropy)


if __name__ == "__main__":
  for name in sys.argv[1:] or ["c3"]:
    (ebm if bench.CONFIGS[name].get("ebm") else state_vector)(name)
