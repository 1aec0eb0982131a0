"""CPU tests of the round-2 scheduler / lowering decisions, read off the compiled plan (no GPU):
sparse first forward sweep, gradient-only first gates, the two-schedule comparison."""
import re

import numpy as np

from oracle import qhbm_oracle as orc
import helpers as hp


def _launches(text):
  return [(int(m.group(1)), int(m.group(2), 16)) for m in re.finditer(r"^launch (\d+) flags=0x([0-9a-f]+)", text, re.M)]


def _passes(text):
  return re.findall(r"^\s+pass (\d+) regbits=\[([0-9 ]+)\] ngrad=(-?\d+) ops:(.*)$", text, re.M)


LF_SPARSE_OUT, LF_SPARSE_IN = 1 << 8, 1 << 9


def test_sparse_first_forward_sweep_is_flagged_only_when_a_second_sweep_follows():
  gates, names = orc.hea_circuit(16, 2)
  flags = dict(_launches(hp.dump_plan(gates, 16, len(names), [orc.xxz_ring(16)], True)))
  assert flags[0] & LF_SPARSE_OUT and not flags[0] & LF_SPARSE_IN
  assert flags[1] & LF_SPARSE_IN and not flags[1] & LF_SPARSE_OUT
  assert not any(f & (LF_SPARSE_OUT | LF_SPARSE_IN) for i, f in flags.items() if i >= 2)
  # a state that fits one tile has one launch and nothing to skip
  gates, names = orc.hea_circuit(12, 2)
  flags = dict(_launches(hp.dump_plan(gates, 12, len(names), [orc.tfim_ring(12)], True)))
  assert len(flags) == 1 and not flags[0] & (LF_SPARSE_OUT | LF_SPARSE_IN)


def test_switch_restores_dense_first_sweep(monkeypatch):
  monkeypatch.setenv("QHBM_NO_SPARSE_INIT", "1")
  gates, names = orc.hea_circuit(16, 2)
  flags = dict(_launches(hp.dump_plan(gates, 16, len(names), [orc.xxz_ring(16)], False)))
  assert not any(f & (LF_SPARSE_OUT | LF_SPARSE_IN) for f in flags.values())


def test_deferring_first_gate_gradients_saves_a_pass_on_the_headline_circuit(monkeypatch):
  gates, names = orc.hea_circuit(16, 2)
  ops = [orc.xxz_ring(16)]
  chosen = hp.dump_plan(gates, 16, len(names), ops, True)
  monkeypatch.setenv("QHBM_NO_TAIL_DEFER", "1")
  plain = hp.dump_plan(gates, 16, len(names), ops, True)
  assert len(_launches(chosen)) == len(_launches(plain)) == 4
  assert len(_passes(chosen)) == len(_passes(plain)) - 1
  # the last backward sweep ends with a pass of gradient-only rotations (no un-application: XROTM only)
  grad_only = [p for p in _passes(chosen) if int(p[2]) > 0 and p[3].strip() == "XROTM=1"]
  assert grad_only, chosen


def test_plan_comparison_never_pays_an_extra_sweep(monkeypatch):
  """HEA followed by an inverse HEA (the QMHL term): deferral would need a third backward sweep, so the
  comparison keeps the plain schedule."""
  n = 16
  g1, names1 = orc.hea_circuit(n, 2, "q")
  g2, names2 = orc.hea_circuit(n, 2, "m")
  gates = orc.concat_circuits(g1, len(names1), orc.inverse_circuit(g2))
  nsym = len(names1) + len(names2)
  ops = orc.kobe_shards(n, 2)
  chosen = hp.dump_plan(gates, n, nsym, ops, True)
  monkeypatch.setenv("QHBM_NO_TAIL_DEFER", "1")
  plain = hp.dump_plan(gates, n, nsym, ops, True)
  assert len(_launches(chosen)) <= len(_launches(plain))
  assert len(_passes(chosen)) <= len(_passes(plain))


def test_first_gates_are_not_unapplied_but_still_differentiated():
  """Gradient of the first-layer rotations against the oracle on a multi-tile plan (their un-application is
  skipped; QHBM_NO_TAIL_SKIP would restore it): covered numerically by the device-program emulator."""
  rng = np.random.default_rng(12)
  n = 11
  gates, names = orc.hea_circuit(n, 1)   # ONE layer: every rotation is a first gate
  phi = rng.uniform(-1, 1, len(names)).astype(np.float32)
  dg = rng.uniform(-1, 1, 1).astype(np.float32)
  ops = [orc.xxz_ring(n)]
  e, g, _, _ = hp.verify_run(gates, n, len(names), ops, phi, 77, dg, True, 9, 4, 0)
  e_ref, g_ref = orc.adjoint_gradient(gates, n, phi, 77, ops, dg, "exact")
  np.testing.assert_allclose(e, e_ref, atol=2e-4)
  np.testing.assert_allclose(g, g_ref, atol=4e-4)
