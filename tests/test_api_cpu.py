"""CPU tests of the host-side mirror of the reference interface (no kernels are called):
circuit construction and lowering, symbol ordering, error behaviour, energy layers.
Each test names the reference test it mirrors (file:line under /root/reference/tests)."""
import itertools
import math
import os

import numpy as np
import pytest
import torch

from oracle import qhbm_oracle as orc
from qhbmlib import _native as nat
from qhbmlib import architectures as arch
from qhbmlib import circuits as cq
from qhbmlib import models
from qhbmlib.models import energy_utils


def test_c_abi_library_exports_every_declared_symbol():
  """The shared library loads and exports every function include/qhbm_b200.h declares."""
  import re
  header = open(os.path.join(os.path.dirname(__file__), "..", "include", "qhbm_b200.h")).read()
  declared = set(re.findall(r"\b(qhbm_[a-z0-9_]+)\s*\(", header))
  assert declared == set(nat.SIGNATURES), declared ^ set(nat.SIGNATURES)
  lib = nat.lib()
  for name in declared:
    assert hasattr(lib, name)
  assert lib.qhbm_version() >= 100


def test_header_is_plain_c99_and_links_from_c(tmp_path):
  """The drop-in boundary is a C ABI: include/qhbm_b200.h must compile as strict C99 (no C++ or torch types in
  the signatures) and a C program must link against the shared library and call into it."""
  import shutil
  import subprocess
  if shutil.which("gcc") is None:
    pytest.skip("no gcc")
  root = os.path.join(os.path.dirname(__file__), "..")
  src = tmp_path / "abi.c"
  src.write_text(
      '#include <stdio.h>\n#include "qhbm_b200.h"\n'
      "int main(void) {\n"
      "  qhbm_gate_t g; qhbm_pauli_term_t t; qhbm_energy_desc_t e; qhbm_comm_t* c = 0; qhbm_plan_t* p = 0;\n"
      "  (void)g; (void)t; (void)e; (void)c; (void)p;\n"
      "  if (qhbm_version() < 100) return 2;\n"
      "  /* an invalid call must come back as a status + message, not crash */\n"
      "  if (qhbm_allreduce(0, 0, 4, QHBM_F32, 0) == 0) return 3;\n"
      '  printf("%s\\n", qhbm_last_error());\n'
      "  return 0;\n}\n")
  exe = tmp_path / "abi"
  lib_dir = os.path.dirname(nat.LIB_PATH)
  subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(root, "include"),
                         str(src), "-o", str(exe), "-L", lib_dir, "-l:libqhbm_b200.so", f"-Wl,-rpath,{lib_dir}"])
  out = subprocess.run([str(exe)], capture_output=True, text=True)
  assert out.returncode == 0 and "null communicator" in out.stdout, (out.returncode, out.stdout, out.stderr)


def test_collective_entry_points_bind_nccl_at_run_time_and_validate():
  """qhbm_comm_unique_id works without a GPU (it only needs libnccl.so.2); bad arguments come back as status 1
  with a message, before NCCL is touched."""
  import ctypes
  lib = nat.lib()
  a, b = np.zeros(nat.COMM_ID_BYTES, np.uint8), np.zeros(nat.COMM_ID_BYTES, np.uint8)
  nat.check(lib.qhbm_comm_unique_id(a.ctypes.data, a.size))
  nat.check(lib.qhbm_comm_unique_id(b.ctypes.data, b.size))
  assert a.any() and (a != b).any()
  assert lib.qhbm_comm_unique_id(a.ctypes.data, 64) == 1
  assert b"QHBM_COMM_ID_BYTES" in lib.qhbm_last_error()
  handle = ctypes.c_void_p()
  assert lib.qhbm_comm_create(a.ctypes.data, 2, 2, ctypes.byref(handle)) == 1
  assert b"rank out of range" in lib.qhbm_last_error()
  assert lib.qhbm_allreduce(None, None, 4, 0, None) == 1
  assert b"null communicator" in lib.qhbm_last_error()
  lib.qhbm_comm_destroy(None)  # a null handle is ignored


def test_no_cuda_device_fails_loudly():
  if torch.cuda.is_available():
    pytest.skip("needs a machine without a GPU")
  from qhbmlib import engine
  gates, names = orc.hea_circuit(3, 1)
  terms, offs = engine.terms_from_pauli_sums([orc.tfim_ring(3)], 3)
  with pytest.raises(nat.NativeError, match="no CPU fallback"):
    engine.ExpectationPlan(gates, 3, len(names), terms, offs)
  with pytest.raises(TypeError, match="no CPU path"):
    from qhbmlib import utils
    utils.unique_bitstrings_with_counts(torch.zeros((2, 2), dtype=torch.int8))


def test_hea_gate_table_matches_oracle_builder():
  """tests/test_util.py:25-67: same gates, same lexicographic symbol order."""
  for n, layers in [(1, 2), (4, 2), (12, 2)]:
    qubits = cq.GridQubit.rect(1, n)
    circ = arch.get_hardware_efficient_model_unitary(qubits, layers, "q")
    qc = models.DirectQuantumCircuit(circ)
    ref_gates, ref_names = orc.hea_circuit(n, layers, "q")
    assert qc.symbol_names == ref_names
    table = qc.gate_table()
    for f in ("type", "q0", "q1", "nparams", "sym", "scalar", "cnst", "gshift"):
      np.testing.assert_array_equal(table[f], ref_gates[f])


def test_bit_injector_symbol_order_quirk():
  """models/circuit.py:59-63 sorts "bit_circuit_bit_{k}" as strings."""
  for n in (3, 10, 12, 20):
    qc = models.DirectQuantumCircuit(arch.get_hardware_efficient_model_unitary(cq.GridQubit.rect(1, n), 1, "a"))
    pi = orc.bit_column_to_qubit(n)
    assert qc._bit_shifts == [n - 1 - pi[j] for j in range(n)]


def test_circuit_add_and_pow():
  """tests/models/circuit_test.py:170-240."""
  qubits = cq.GridQubit.rect(1, 3)
  a = models.DirectQuantumCircuit(arch.get_hardware_efficient_model_unitary(qubits, 1, "a"), name="a")
  b = models.DirectQuantumCircuit(arch.get_hardware_efficient_model_unitary(qubits[:2], 1, "b"), name="b")
  s = a + b
  assert s.symbol_names == a.symbol_names + b.symbol_names
  assert s.qubits == sorted(qubits)
  assert len(s.pqc) == len(a.pqc) + len(b.pqc)
  assert {id(p) for p in s.parameters()} == {id(p) for p in a.parameters()} | {id(p) for p in b.parameters()}
  np.testing.assert_allclose(s.symbol_values.detach(), torch.cat([a.symbol_values, b.symbol_values]).detach())
  with pytest.raises(ValueError, match="must not have symbols in common"):
    _ = a + a
  with pytest.raises(TypeError):
    _ = a + 1
  inv = a**-1
  assert inv.symbol_names == a.symbol_names
  assert [id(p) for p in inv.parameters()] == [id(p) for p in a.parameters()]
  tab, itab = a.gate_table(), inv.gate_table()
  np.testing.assert_array_equal(itab["type"], tab["type"][::-1])
  np.testing.assert_allclose(itab["scalar"][:, 0], -tab["scalar"][::-1, 0])
  with pytest.raises(ValueError, match="Only the inverse"):
    _ = a**2


def test_inverse_gate_table_matches_oracle_inverse():
  qubits = cq.GridQubit.rect(1, 4)
  s = cq.symbols("a b c d")
  circ = cq.Circuit(cq.rx(s[0])(qubits[0]), cq.ISWAP(qubits[0], qubits[1])**s[1],
                    cq.PhasedXPowGate(0.3, s[2], -0.5)(qubits[2]), cq.FSimGate(s[3], 0.7)(qubits[2], qubits[3]),
                    cq.XXPowGate(exponent=0.5 * s[0], global_shift=-0.5)(qubits[3], qubits[0]))
  names = sorted(cq.circuit_symbols(circ))
  fwd = cq.gate_table(circ, qubits, names)
  inv = cq.gate_table(circ**-1, qubits, names)
  ref = orc.inverse_circuit(fwd.astype(orc.GATE_DTYPE))
  for f in ("type", "q0", "q1", "sym", "gshift"):
    np.testing.assert_array_equal(inv[f], ref[f])
  np.testing.assert_allclose(inv["scalar"], ref["scalar"])
  np.testing.assert_allclose(inv["cnst"], ref["cnst"])
  np.testing.assert_allclose(fwd["scalar"][0, 0], 1 / math.pi, rtol=1e-7)
  assert fwd["gshift"][0] == -0.5


def test_sympy_symbols_are_accepted():
  sympy = pytest.importorskip("sympy")
  q = cq.GridQubit(0, 0)
  p = sympy.Symbol("p")
  circ = cq.Circuit(cq.X(q)**p, cq.Z(q)**(2 * p), cq.rx(p)(q))
  tab = cq.gate_table(circ, [q], ["p"])
  np.testing.assert_allclose(tab["scalar"][:, 0], [1.0, 2.0, 1 / math.pi], rtol=1e-6)
  with pytest.raises(ValueError, match="at most one symbol"):
    cq.as_param(p * sympy.Symbol("r"))


def test_pauli_algebra_and_tables():
  q = cq.GridQubit.rect(1, 3)
  h = cq.PauliSum()
  h -= 2.0 * cq.X(q[0])
  h += cq.Z(q[0]) * cq.Z(q[1])
  h += cq.Z(q[0]) * cq.Z(q[1])
  assert len(h.terms) == 2
  assert (cq.X(q[0]) * cq.Y(q[0])).coefficient == 1j and (cq.X(q[0]) * cq.Y(q[0])).paulis == {q[0]: "Z"}
  t, o = cq.convert_to_tensor([h, cq.PauliSum.from_pauli_strings(cq.Y(q[2]))]).tables(q)
  assert list(o) == [0, 2, 3]
  rows = {(float(r["coeff"]), int(r["xmask"]), int(r["zmask"])) for r in t}
  assert rows == {(-2.0, 4, 0), (2.0, 0, 6), (1.0, 1, 1)}
  assert arch.tfim_ring(q) == arch.tfim_ring(q)
  assert len(arch.xxz_ring(cq.GridQubit.rect(1, 16)).terms) == 48


def test_hamiltonian_size_mismatch_raises():
  """tests/models/hamiltonian_test.py:70-81."""
  qubits = cq.GridQubit.rect(1, 3)
  circ = models.DirectQuantumCircuit(arch.get_hardware_efficient_model_unitary(qubits, 1, "x"))
  with pytest.raises(ValueError, match="same number of bits"):
    models.Hamiltonian(models.BernoulliEnergy([0, 1]), circ)
  h = models.Hamiltonian(models.KOBE([0, 1, 2], 2), circ)
  assert h.operator_shards.shape == (6,)
  assert isinstance(h.circuit_dagger, models.QuantumCircuit)
  h2 = models.Hamiltonian(models.BitstringEnergy([0, 1, 2], [torch.nn.Linear(3, 1)]), circ)
  assert h2.operator_shards is None


def test_check_helpers():
  """tests/models/energy_utils_test.py:26-45."""
  assert energy_utils.check_bits([1, 5, 7]) == [1, 5, 7]
  with pytest.raises(ValueError, match="must be unique"):
    energy_utils.check_bits([1, 1])
  assert energy_utils.check_order(3) == 3
  with pytest.raises(TypeError, match="must be an integer"):
    energy_utils.check_order("a")
  with pytest.raises(ValueError, match="greater than zero"):
    energy_utils.check_order(0)


def test_parity_layer_golden():
  """tests/models/energy_utils_test.py:86-110."""
  layer = models.Parity([1, 2, 3, 4], 3)
  assert layer.indices == [[0], [1], [2], [3], [0, 1], [0, 2], [0, 3], [1, 2], [1, 3], [2, 3], [0, 1, 2],
                           [0, 1, 3], [0, 2, 3], [1, 2, 3]]
  assert layer.num_terms == 14
  out = layer(torch.tensor([[-1, 1, -1, -1]]))
  np.testing.assert_array_equal(out.numpy(), [[-1, 1, -1, -1] + [-1, 1, 1, -1, -1, 1] + [1, 1, -1, 1]])


def test_bernoulli_energy_golden():
  """tests/models/energy_test.py:111-182."""
  b = models.BernoulliEnergy([1, 2, 3])
  v = torch.tensor([1.0, 1.7, -2.8])
  b.set_weights([v])
  bits = torch.tensor([[0, 0, 0], [1, 0, 0], [0, 1, 1]], dtype=torch.int8)
  np.testing.assert_allclose(b.logits.detach(), 2 * v)
  e = b(bits)
  np.testing.assert_allclose(e.detach(), [v[0] + v[1] + v[2], -v[0] + v[1] + v[2], v[0] - v[1] - v[2]], rtol=1e-6)
  jac = torch.autograd.functional.jacobian(lambda k: torch.sum((1 - 2 * bits).float() * k, -1), v)
  np.testing.assert_array_equal(jac.numpy(), (1 - 2 * bits).numpy())
  assert [s.terms[0].paulis for s in b.operator_shards(cq.GridQubit.rect(1, 3))] == [
      {q: "Z"} for q in cq.GridQubit.rect(1, 3)]


def test_kobe_energy_golden():
  """tests/models/energy_test.py:233-266."""
  k = models.KOBE([0, 1], 2)
  k.set_weights([torch.tensor([1.5, 2.7, -4.0])])
  e = k(torch.tensor([[0, 0], [0, 1], [1, 0], [1, 1]], dtype=torch.int8))
  np.testing.assert_allclose(e.detach(), [0.2, 2.8, 5.2, -8.2], rtol=1e-6)
  q = cq.GridQubit.rect(1, 3)
  shards = models.KOBE([0, 1, 2], 2).operator_shards(q)
  assert [s.terms[0].paulis for s in shards] == [{q[0]: "Z"}, {q[1]: "Z"}, {q[2]: "Z"}, {q[0]: "Z", q[1]: "Z"},
                                                 {q[0]: "Z", q[2]: "Z"}, {q[1]: "Z", q[2]: "Z"}]
  kk = models.KOBE(list(range(5)), 3)
  bits = torch.tensor(list(itertools.product([0, 1], repeat=5)), dtype=torch.int8)
  th = kk.post_process[0].kernel.detach().numpy()
  np.testing.assert_allclose(kk(bits).detach().numpy(), orc.kobe_energy(bits.numpy(), 3, th), rtol=1e-5, atol=1e-6)


def test_seed_helpers():
  from qhbmlib.inference import ebm
  s = ebm.sanitize_seed([5, 6])
  assert s.tolist() == [5, 6]
  a, b = ebm.split_seed(s)
  a2, _ = ebm.split_seed(s)
  assert a.tolist() == a2.tolist() and a.tolist() != b.tolist() and a.tolist() != s.tolist()
  assert ebm.sanitize_seed(None).tolist() != ebm.sanitize_seed(None).tolist()


def test_qaia_structure():
  """reference circuit.py:211-292: symbol naming and tied parameters."""
  q = cq.GridQubit.rect(1, 2)
  quantum = [cq.PauliSum.from_pauli_strings(cq.X(q[0])), cq.PauliSum.from_pauli_strings(cq.X(q[1]))]
  classical = [cq.PauliSum.from_pauli_strings(cq.Z(q[0])), cq.PauliSum.from_pauli_strings(cq.Z(q[0]) * cq.Z(q[1]))]
  qaia = models.QAIA(quantum, classical, 3)
  assert qaia.symbol_names[:4] == ["gamma_0_0", "gamma_0_1", "eta_0_0", "eta_0_1"]
  assert qaia.symbol_values.shape == (12,)
  etas, thetas, gammas = qaia.value_layers_inputs[0]
  exp = torch.cat([etas.unsqueeze(1) * thetas.unsqueeze(0), gammas], 1).reshape(-1)
  np.testing.assert_allclose(qaia.symbol_values.detach(), exp.detach())


# ---------------------------------------------------------------- shot-based inference: host logic
def test_parameter_shift_occurrence_table():
  """Every shiftable symbolic exponent becomes its own symbol whose value is scalar*phi+const;
  three-eigenvalue gates and phase exponents are reported as blocked."""
  import torch
  from qhbmlib import circuits as cq
  from qhbmlib.inference import qnn
  q = cq.GridQubit.rect(1, 3)
  a, b = cq.Symbol("a"), cq.Symbol("b")
  circuit = cq.Circuit(cq.X(q[0])**(a * 2.0), cq.rx(0.3).on(q[1]), cq.CZ(q[0], q[1])**(b * -0.5 + 0.25),
                       cq.Y(q[2])**a, cq.H(q[1])**0.5)
  names = ["a", "b"]
  table = cq.gate_table(circuit, q, names)
  occ = qnn._Occurrences(table, 2)
  assert occ.total_symbols == 5 and occ.blocked == []
  assert occ.sym.tolist() == [0, 1, 0]
  np.testing.assert_allclose(occ.scalar.numpy(), [2.0, -0.5, 1.0])
  np.testing.assert_allclose(occ.const.numpy(), [0.0, 0.25, 0.0])
  assert [int(r["sym"][0]) for r in occ.table] == [2, -1, 3, 4, -1]
  assert all(float(r["scalar"][0]) == 1.0 and float(r["cnst"][0]) == 0.0 for r in occ.table[[0, 2, 3]])
  assert int(table[0]["sym"][0]) == 0  # the circuit's own table is untouched
  vals = occ.values(torch.tensor([0.3, -0.8]))
  np.testing.assert_allclose(vals.numpy(), [0.3, -0.8, 0.6, 0.65, 0.3], rtol=1e-6)
  blocked = cq.gate_table(cq.Circuit(cq.ISWAP(q[0], q[1])**b, cq.X(q[0])**a), q, names)
  occ = qnn._Occurrences(blocked, 2)
  assert occ.blocked == [1] and occ.sym.tolist() == [0]


def test_sampled_split_terms():
  """PauliSums -> unit Pauli strings + mixing matrix + identity offsets."""
  from qhbmlib import circuits as cq
  from qhbmlib.inference import qnn
  q = cq.GridQubit.rect(1, 2)
  ops = cq.convert_to_tensor([
      cq.PauliSum.from_pauli_strings(2.0 * cq.X(q[0])) + cq.PauliSum.from_pauli_strings(-0.5 * cq.Z(q[0]) * cq.Z(q[1])) +
      cq.PauliSum.from_pauli_strings(cq.PauliString(1.5, {})),
      cq.PauliSum.from_pauli_strings(3.0 * cq.Y(q[1])),
  ])
  term_ops, mix, offsets = qnn.SampledQuantumInference._split_terms(ops, q)
  assert len(term_ops) == 3
  assert all(len(s.terms) == 1 and s.terms[0].coefficient == 1.0 for s in term_ops.pauli_sums)
  np.testing.assert_allclose(mix.numpy(), [[2.0, 0.0], [-0.5, 0.0], [0.0, 3.0]])
  np.testing.assert_allclose(offsets.numpy(), [1.5, 0.0])
  only_identity = cq.convert_to_tensor([cq.PauliSum.from_pauli_strings(cq.PauliString(0.7, {}))])
  term_ops, mix, offsets = qnn.SampledQuantumInference._split_terms(only_identity, q)
  assert len(term_ops) == 1 and float(mix.abs().sum()) == 0.0 and offsets.tolist() == pytest.approx([0.7])


def test_variables_updated_sees_writes_through_dot_data():
  """`p.data.add_()` changes neither the storage pointer nor the version counter; the preface
  must still notice it (the reference compares values on every call, ebm.py:125-140)."""
  from qhbmlib import inference
  energy = models.BernoulliEnergy([0, 1, 2], energy_utils.Constant(0.3) if hasattr(energy_utils, "Constant") else None)
  inf = inference.BernoulliEnergyInference(energy, 10, initial_seed=1)
  kernel = energy.post_process[0].kernel
  with torch.no_grad():
    kernel.copy_(torch.tensor([0.3, -0.2, 0.5]))
  with torch.no_grad():  # (with gradients the log-partition draws samples: CUDA only)
    lp0 = float(inf.log_partition())
  assert not inf.variables_updated
  ptr, version = kernel.data_ptr(), kernel._version
  kernel.data.add_(1.0)
  assert (kernel.data_ptr(), kernel._version) == (ptr, version)   # the cheap test would have missed it
  assert inf.variables_updated
  with torch.no_grad():
    lp1 = float(inf.log_partition())
  t = np.array([1.3, 0.8, 1.5])
  np.testing.assert_allclose(lp1, np.sum(np.log(2 * np.cosh(t))), rtol=1e-6)
  assert abs(lp1 - lp0) > 0.5


def test_plan_cache_lru_is_bounded_and_content_keyed():
  from qhbmlib.inference import qnn
  lru = qnn._LRU(3)
  made = []
  for key in ["a", "b", "a", "c", "d", "a", "b"]:
    lru.get(key, lambda k=key: made.append(k) or k.upper())
  assert made == ["a", "b", "c", "d", "b"] and len(lru) == 3   # "b" was evicted by "d", "a" stayed hot
  qubits = cq.GridQubit.rect(1, 3)
  h1 = cq.convert_to_tensor([cq.Z(qubits[0]) * cq.Z(qubits[1]) + 0.5 * cq.X(qubits[2])])
  h2 = cq.convert_to_tensor([cq.Z(qubits[0]) * cq.Z(qubits[1]) + 0.5 * cq.X(qubits[2])])
  h3 = cq.convert_to_tensor([cq.Z(qubits[0]) * cq.Z(qubits[1]) + 0.25 * cq.X(qubits[2])])
  assert h1 is not h2 and h1.tables_digest(qubits) == h2.tables_digest(qubits) != h3.tables_digest(qubits)


def test_log_partition_gradient_is_drawn_lazily_in_backward():
  """reference ebm.py:331-343, 396-415: the log-partition gradient estimator samples inside grad_fn, i.e. only
  when the gradient is taken.  `_LogPartitionGrad` on a stand-in owner: no surrogate evaluation in forward or
  when the result is detached (vqt), one per backward, gradient = upstream * d surrogate / d theta."""
  from qhbmlib.inference import ebm

  class Owner:
    calls = 0

    def __init__(self):
      self.theta = torch.nn.Parameter(torch.tensor([0.5, -1.0, 2.0]))
      self.unused = torch.nn.Parameter(torch.tensor([1.0]))

    def _log_partition_surrogate(self, sharded):
      assert sharded is False and torch.is_grad_enabled()
      Owner.calls += 1
      return -(self.theta * torch.tensor([1.0, 2.0, 3.0])).sum()

  owner = Owner()
  value = torch.tensor(1.25)
  out = ebm._LogPartitionGrad.apply(owner, value, False, owner.theta, owner.unused)
  assert float(out) == 1.25 and out.requires_grad and Owner.calls == 0
  _ = out.detach() * 3.0  # vqt's use: nothing is sampled
  assert Owner.calls == 0
  (2.0 * out).backward()
  assert Owner.calls == 1
  np.testing.assert_allclose(owner.theta.grad.numpy(), [-2.0, -4.0, -6.0])
  assert owner.unused.grad is None
