"""Single-process emulation of the rank-sharded categorical draw (one GPU): the full draw against the union
of two half-range draws with mass intervals, as AnalyticEnergyInference._draw_keys does over NCCL."""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "qhbm-library_b200")):
  sys.path.insert(0, p)
from qhbmlib import engine  # noqa: E402

dev = torch.device("cuda", 0)
for n, ns in ((12, 200000), (14, 1000000), (20, 1000000)):
  g = torch.Generator(device="cpu").manual_seed(n)
  logits = (torch.randn(1 << n, generator=g) * 2.0).float().to(dev)
  gmax = float(logits.double().max())
  full = engine.CategoricalSampler(logits, given_max=gmax)
  seed = (123, 456)
  ref = full.draw(ns, seed)
  for world in (2, 3, 8):
    bounds = [(r * (1 << n)) // world for r in range(world + 1)]
    samplers = [engine.CategoricalSampler(logits[bounds[r]:bounds[r + 1]].contiguous(), given_max=gmax) for r in range(world)]
    cum = [0.0]
    for s in samplers:
      cum.append(cum[-1] + float(s.local_mass().item()))
    out = torch.zeros((ns,), dtype=torch.int64, device=dev)
    hits = torch.zeros((ns,), dtype=torch.int64, device=dev)
    for r, s in enumerate(samplers):
      o = torch.full((ns,), -1, dtype=torch.int64, device=dev)
      s.draw(ns, seed, row_offset=bounds[r], mass_interval=(cum[r], cum[r + 1] if r + 1 < world else math.inf, cum[-1]), out=o)
      hits += (o >= 0).long()
      out += torch.where(o >= 0, o, torch.zeros_like(o))
    bad = (out != ref).nonzero().flatten()
    print(f"n={n} world={world} samples={ns}: mismatches {bad.numel()}, claimed-by-one {(hits == 1).all().item()}",
          "total full", float(full.local_mass().item()), "sum shards", cum[-1])
    for k in bad[:5].tolist():
      print("   sample", k, "full", int(ref[k]), "sharded", int(out[k]), "logit full", float(logits[ref[k]]),
            "logit sharded", float(logits[out[k]]))
