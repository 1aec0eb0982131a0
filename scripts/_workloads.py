"""Gate / Pauli tables of the development scripts, from the product's own builders (qhbmlib.architectures,
qhbmlib.circuits) -- nothing under scripts/ imports oracle/ (test infrastructure)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "qhbm-library_b200")):
  if p not in sys.path:
    sys.path.insert(0, p)


def hea_tables(n, layers, ham="xxz"):
  """(gate table, symbol count, term table, term offsets) of HEA(n, layers) measured on `ham`:
  "xxz" / "tfim" ring, or "kobe" = the Z-string shards of a second-order KOBE energy."""
  from qhbmlib import architectures as arch
  from qhbmlib import circuits as cq
  from qhbmlib import models
  qubits = cq.GridQubit.rect(1, n)
  circuit = arch.get_hardware_efficient_model_unitary(qubits, layers, "q")
  names = sorted(cq.circuit_symbols(circuit))
  if ham == "kobe":
    ops = cq.convert_to_tensor(models.KOBE(list(range(n)), 2).operator_shards(qubits))
  else:
    ops = cq.convert_to_tensor([arch.xxz_ring(qubits) if ham == "xxz" else arch.tfim_ring(qubits)])
  terms, offs = ops.tables(qubits)
  return cq.gate_table(circuit, qubits, names), len(names), terms, offs
