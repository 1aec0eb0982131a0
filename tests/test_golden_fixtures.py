"""Committed golden vectors (tests/golden/qhbm_golden.npz, made by tests/golden/make_golden.py).
CPU: the oracle still reproduces them.  GPU: the CUDA path hits the same numbers through the C ABI."""
import os

import numpy as np
import pytest

from oracle import qhbm_oracle as orc
import helpers as hp

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "qhbm_golden.npz"))
CASES = [("c1_tfim4", 4, 2, "tfim"), ("tfim8", 8, 2, "tfim"), ("xxz10", 10, 3, "xxz"), ("c2_tfim12", 12, 2, "tfim"),
         ("c3_xxz16", 16, 2, "xxz")]


def _problem(name, n, layers, ham):
  gates, names = orc.hea_circuit(n, layers)
  ops = [orc.tfim_ring(n) if ham == "tfim" else orc.xxz_ring(n)] + orc.kobe_shards(n, 2)[:3]
  return gates, names, ops, GOLD[f"{name}/phi"], GOLD[f"{name}/basis"], GOLD[f"{name}/dgrad"]


@pytest.mark.parametrize("name,n,layers,ham", CASES[:3])
def test_oracle_reproduces_golden(name, n, layers, ham):
  gates, names, ops, phi, basis, dg = _problem(name, n, layers, ham)
  for mode in ("exact", "tfq_fd"):
    e, g = orc.batch_expectation_and_gradient(gates, n, phi, basis, ops, dg, mode)
    np.testing.assert_allclose(e, GOLD[f"{name}/exp"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(g, GOLD[f"{name}/{mode}/grad"], rtol=1e-10, atol=1e-12)


def test_oracle_reproduces_golden_ebm():
  en = orc.kobe_energy(orc.all_bitstrings(10), 2, GOLD["kobe10/theta"])
  np.testing.assert_allclose(en, GOLD["kobe10/energies"], rtol=1e-12)
  np.testing.assert_allclose([orc.analytic_log_partition(en), orc.analytic_entropy(en)], GOLD["kobe10/logz_entropy"],
                             rtol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("name,n,layers,ham", CASES)
@pytest.mark.parametrize("mode", ["exact", "tfq_fd"])
def test_cuda_path_hits_golden(name, n, layers, ham, mode):
  import torch
  from qhbmlib import engine
  gates, names, ops, phi, basis, dg = _problem(name, n, layers, ham)
  terms, offs = hp.ops_to_tables(ops, n)
  plan = engine.ExpectationPlan(gates, n, len(names), terms, offs, True)
  e, g = plan.forward_adjoint(torch.tensor(basis.astype(np.int64), device="cuda"), torch.tensor(phi, device="cuda"),
                              torch.tensor(dg, device="cuda"), per_state=True, grad_mode=mode)
  scale = max(sum(abs(c) for c, _ in op) for op in ops)
  np.testing.assert_allclose(e.cpu().numpy(), GOLD[f"{name}/exp"], rtol=1e-5, atol=1e-5 * scale)
  gref = GOLD[f"{name}/{mode}/grad"]
  np.testing.assert_allclose(g.cpu().numpy(), gref, rtol=1e-5, atol=3e-5 * np.abs(gref).max())


@pytest.mark.gpu
def test_cuda_ebm_hits_golden():
  import torch
  from qhbmlib import _native as nat
  from qhbmlib import engine
  n = 10
  masks = np.array([sum(1 << (n - 1 - q) for q in c) for c in orc.parity_indices(n, 2)], dtype=np.int32)
  d = engine.EnergyDescriptor(nat.ENERGY_KOBE, n, torch.tensor(masks, device="cuda"),
                              torch.tensor(GOLD["kobe10/theta"], device="cuda"))
  logits, stats = d.sweep(0, 1 << n)
  np.testing.assert_allclose(-logits.cpu().numpy(), GOLD["kobe10/energies"], rtol=1e-5, atol=1e-5)
  m, s, t = stats.cpu().numpy()
  np.testing.assert_allclose([m + np.log(s), m + np.log(s) - t / s], GOLD["kobe10/logz_entropy"], rtol=1e-6)
