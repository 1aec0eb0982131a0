// plan.h -- host-side compiler: gate table + Pauli sums -> sweep/pass/op program.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/qhbm_b200.h"
#include "program.h"

namespace qhbm {

struct CircuitIR {
  int n_qubits = 0;
  int n_symbols = 0;
  std::vector<qhbm_gate_t> gates;
};

struct OpsIR {
  int n_qubits = 0;
  std::vector<qhbm_pauli_term_t> terms;
  std::vector<int32_t> offsets;  // n_ops + 1
  int n_ops() const { return (int)offsets.size() - 1; }
};

struct HostPlan {
  int n = 0;       // circuit qubits
  int n_eff = 0;   // simulated qubits (idle high qubits pad tiny circuits)
  int T = 0;       // tile qubits
  int K = 0;       // register qubits
  int P = 0;       // symbols
  int O = 0;       // observables
  bool grad = false;
  int n_sweeps_fwd = 0, n_sweeps_bwd = 0;

  std::vector<qhbm_gate_t> gates;
  std::vector<DevPass> passes;
  std::vector<DevOp> ops;
  std::vector<int32_t> gsym;
  std::vector<PrepJob> jobs;
  std::vector<int32_t> lists;
  int32_t ncoef = 0;
  int32_t phase_coef = -1;  // coefficient slot of the dropped global phase (debug state output)
  std::vector<LaunchDesc> launches;       // forward (+ expectation + backward) program
  int n_fwd_launches = 0;                 // launches[0..n_fwd_launches) leave psi = U|basis>

  std::vector<DevTerm> terms;
  std::vector<DevTermGroup> groups;
  std::vector<DevOpRange> opranges;
  std::vector<DevDiagTerm> dterms;        // diagonal terms routed through the WHT path (may be empty)

  // Device program (lower_device_program): the passes / ops the kernels execute.  Same schedule as
  // `passes` / `ops`; gradient passes additionally carry their reduction tasks, scratch-unit numbers
  // instead of gradient slots, and the flush windows of the gradient descriptors.
  std::vector<DevPass> dev_passes;
  std::vector<PackedOp> dev_ops;
  std::vector<DevGradDesc> gdescs;
  bool lean = false;  // no general-matrix op anywhere: the kernels' GEN = false instantiations run it

  int tiles() const { return 1 << (n_eff - T); }
};

// Throws std::runtime_error with a readable message on invalid input.
void validate_circuit(const CircuitIR& c);
void validate_ops(const OpsIR& o);
HostPlan compile_plan(const CircuitIR& c, const OpsIR& o, bool with_gradient, int tile_qubits,
                      int reg_qubits);

}  // namespace qhbm
