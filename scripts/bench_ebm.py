"""EBM-side kernels at BASELINE config 5 scale: 2^24-row sweep with a dense-stack (MLP) energy and with
KOBE-2, logsumexp/entropy statistics, 1e6 categorical samples, first-occurrence unique.  Prints one JSON
line per kernel with achieved FLOP/s or GB/s (device time, CUDA events, best of 5)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "qhbm-library_b200")):
  sys.path.insert(0, p)
import numpy as np
import torch
from qhbmlib import _native as nat
from qhbmlib import engine
from qhbmlib import models


def timeit(f, reps=5):
  f()
  torch.cuda.synchronize()
  best = 1e9
  for _ in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = f()
    b.record()
    torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b))
  return best, out


def main():
  n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
  rows = 1 << n
  rng = np.random.default_rng(4)
  widths = [n, 64, 64, 1]
  layers = []
  for l in range(3):
    lim = np.sqrt(6.0 / (widths[l] + widths[l + 1]))
    layers.append((torch.tensor(rng.uniform(-lim, lim, (widths[l], widths[l + 1])).astype(np.float32), device="cuda"),
                   torch.zeros(widths[l + 1], device="cuda"), ["tanh", "tanh", "linear"][l]))
  mlp = engine.EnergyDescriptor(nat.ENERGY_MLP, n, layers=layers)
  ms, (logits, stats) = timeit(lambda: mlp.sweep(0, rows))
  flops = rows * 2 * sum(widths[l] * widths[l + 1] for l in range(3))
  print(json.dumps({"kernel": "ebm_sweep_kernel (MLP 64-64-1)", "rows": rows, "ms": ms, "rows_per_s": rows / ms * 1e3,
                    "tflops_fp32": flops / ms / 1e9, "frac_of_74.4_nominal": flops / ms / 1e9 / 74.4}))
  parity = models.Parity(list(range(n)), 2)
  masks = torch.tensor(parity.masks(), dtype=torch.int64, device="cuda").to(torch.int32)
  theta = torch.tensor(rng.normal(0, 0.1, len(parity.masks())).astype(np.float32), device="cuda")
  kobe = engine.EnergyDescriptor(nat.ENERGY_KOBE, n, masks, theta)
  ms, _ = timeit(lambda: kobe.sweep(0, rows))
  print(json.dumps({"kernel": "ebm_sweep_kernel (KOBE-2, %d terms)" % len(parity.masks()), "rows": rows, "ms": ms,
                    "rows_per_s": rows / ms * 1e3}))
  n_samples = 1_000_000
  ms, samples = timeit(lambda: engine.categorical_sample(logits, n_samples, (3, 4)))
  print(json.dumps({"kernel": "categorical_sample (max, block sums, scan, sample)", "rows": rows, "samples": n_samples,
                    "ms": ms, "samples_per_s": n_samples / ms * 1e3, "GBps_logits_read": rows * 4 * 2 / ms / 1e6}))
  ms, (uq, idx, cnt) = timeit(lambda: engine.unique_with_counts(samples))
  print(json.dumps({"kernel": "unique_with_counts (first occurrence)", "rows": n_samples, "unique": int(uq.shape[0]),
                    "ms": ms, "rows_per_s": n_samples / ms * 1e3}))
  big = torch.randint(0, 1 << 20, (10_000_000,), device="cuda")
  ms, (uq, idx, cnt) = timeit(lambda: engine.unique_with_counts(big))
  print(json.dumps({"kernel": "unique_with_counts (first occurrence)", "rows": 10_000_000, "unique": int(uq.shape[0]),
                    "ms": ms, "rows_per_s": 1e7 / ms * 1e3, "GBps_keys": 1e7 * 8 / ms / 1e6}))
  m, s, t = stats.tolist()
  print(json.dumps({"logZ": m + np.log(s), "entropy": m + np.log(s) - t / s}))


if __name__ == "__main__":
  main()
