"""One GPU: full-range sweep statistics against merged half-range sweeps and a float64 logsumexp of the logits."""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "qhbm-library_b200")):
  sys.path.insert(0, p)
from qhbmlib import distributed as qd  # noqa: E402
from qhbmlib import inference, models  # noqa: E402
from qhbmlib.inference.ebm import energy_descriptor  # noqa: E402
from qhbmlib.models import energy_utils  # noqa: E402

dev = torch.device("cuda", 0)
for n in (10, 12, 14):
  for seed in (21, 31, 41):
    energy = models.KOBE(list(range(n)), 2, energy_utils.RandomNormal(0.0, 0.3, seed))
    inf = inference.AnalyticEnergyInference(energy, 10, initial_seed=[seed, seed + 1])
    api_logz = float(inf.log_partition())
    dev = inf._logits.device
    desc = energy_descriptor(inf.energy)
    logits, stats = desc.sweep(0, 1 << n, device=dev)
    m, s, t = stats.cpu().tolist()
    print("   api log_partition", repr(api_logz), "api stats", inf._stats.tolist(), "sweep stats", (m, s, t))
    halves = []
    for lo, hi in ((0, 1 << (n - 1)), (1 << (n - 1), 1 << n)):
      l2, st2 = desc.sweep(lo, hi, device=dev)
      halves.append(st2.cpu().tolist())
      if not torch.equal(l2, logits[lo:hi]):
        d = (l2.double() - logits[lo:hi].double()).abs()
        print(f"   logits differ between full and half sweeps on [{lo},{hi}): max abs {float(d.max()):.3e} at row {lo + int(d.argmax())}, "
              f"{int((d > 0).sum())} rows; full {float(logits[lo + int(d.argmax())])!r} half {float(l2[int(d.argmax())])!r}")
    mm, ss, tt = qd.merge_log_stats(halves)
    ref = float(torch.logsumexp(logits.double(), 0))
    print(f"n={n} seed={seed}: full {m + math.log(s)!r} merged {mm + math.log(ss)!r} torch f64 {ref!r}", flush=True)
    # repeat the full sweep a few times: run-to-run stability of the statistics
    vals = set()
    for _ in range(5):
      _, st = desc.sweep(0, 1 << n, device=dev)
      a, b, _c = st.cpu().tolist()
      vals.add(a + math.log(b))
    print("   repeated full sweeps:", sorted(vals))
