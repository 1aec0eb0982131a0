"""The C restatement (oracle/tfq_cpu.c, complex64) against the NumPy oracle (complex128)."""
import numpy as np
import pytest

from oracle import qhbm_oracle as orc
from oracle import tfq_cpu
import helpers as hp


@pytest.mark.parametrize("n,layers", [(3, 2), (6, 2), (10, 1)])
def test_hea_expectation_and_adjoint(n, layers):
  rng = np.random.default_rng(n)
  gates, names = orc.hea_circuit(n, layers)
  phi = rng.uniform(-1, 1, len(names)).astype(np.float32)
  ops = [orc.tfim_ring(n), orc.xxz_ring(n)]
  basis = rng.choice(1 << n, 5, replace=False)
  dg = rng.uniform(-1, 1, (5, 2)).astype(np.float32)
  for mode in ("exact", "tfq_fd"):
    prob = tfq_cpu.Problem(gates, n, phi, ops, mode)
    e_ref, g_ref = orc.batch_expectation_and_gradient(gates, n, phi, basis, ops, dg, mode)
    np.testing.assert_allclose(prob.expectation(basis), e_ref, rtol=1e-5, atol=1e-5 * n)
    e, g = prob.adjoint(basis, dg)
    np.testing.assert_allclose(e, e_ref, rtol=1e-5, atol=1e-5 * n)
    tol = 1e-5 if mode == "exact" else 5e-4  # float32 finite differencing noise (TFQ does the same)
    np.testing.assert_allclose(g, g_ref, rtol=tol, atol=tol * np.abs(g_ref).max())


@pytest.mark.parametrize("seed", range(4))
def test_random_circuits(seed):
  rng = np.random.default_rng(seed)
  n = 5
  gates = hp.random_circuit(n, 30, 6, rng)
  ops = hp.random_ops(n, 3, rng)
  phi = rng.uniform(-1, 1, 6).astype(np.float32)
  basis = rng.choice(1 << n, 4, replace=False)
  dg = rng.uniform(-1, 1, (4, 3)).astype(np.float32)
  prob = tfq_cpu.Problem(gates, n, phi, ops, "exact")
  e_ref, g_ref = orc.batch_expectation_and_gradient(gates, n, phi, basis, ops, dg, "exact")
  e, g = prob.adjoint(basis, dg)
  np.testing.assert_allclose(e, e_ref, rtol=1e-5, atol=2e-5)
  np.testing.assert_allclose(g, g_ref, rtol=1e-4, atol=1e-4 * (np.abs(g_ref).max() + 1))


def test_fusion_block_count_matches_survey():
  """HEA(n, L) fuses into L(n-1) two-qubit blocks."""
  for n, layers in [(4, 2), (8, 3)]:
    gates, names = orc.hea_circuit(n, layers)
    blocks = tfq_cpu.fuse_blocks(gates, np.zeros(len(names)))
    assert len(blocks) == layers * (n - 1)
