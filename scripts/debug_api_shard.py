"""2-rank diagnostic (torchrun): where do the sharded and the single-GPU VQT / QMHL losses part ways?"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "qhbm-library_b200")):
  sys.path.insert(0, p)
from qhbmlib import architectures as arch  # noqa: E402
from qhbmlib import circuits as cq  # noqa: E402
from qhbmlib import data as qdata  # noqa: E402
from qhbmlib import distributed as qd  # noqa: E402
from qhbmlib import engine, inference, models  # noqa: E402
from qhbmlib.models import energy_utils  # noqa: E402

rank = int(os.environ["RANK"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
n, num_samples = 12, 3000
qubits = cq.GridQubit.rect(1, n)


def make_qhbm(tag, seed):
  energy = models.KOBE(list(range(n)), 2, energy_utils.RandomNormal(0.0, 0.3, seed))
  e_inf = inference.AnalyticEnergyInference(energy, num_samples, initial_seed=[seed, seed + 1])
  circ = models.DirectQuantumCircuit(arch.get_hardware_efficient_model_unitary(qubits, 2, tag),
                                     energy_utils.RandomUniform(-1, 1, seed + 2))
  return inference.QHBM(e_inf, inference.AnalyticQuantumInference(circ, grad_mode="exact")), energy, circ


def say(*a):
  if rank == 0:
    print(*a, flush=True)


# 1. the draw
qhbm, energy, circ = make_qhbm("v", 21)
e = qhbm.e_inference
e.entropy()  # triggers _ready_inference
seed = (21, 22)
sh = e._draw_keys(num_samples, seed).clone()
full = e._full_logits()
ref_sampler = engine.CategoricalSampler(full.contiguous(), given_max=float(e._stats[0]))
ref = ref_sampler.draw(num_samples, seed)
bad = (sh != ref).nonzero().flatten()
say("draw: sharded vs full-logits sampler mismatches:", bad.numel(), "of", num_samples, "mass interval", e._mass_interval,
    "full mass", float(ref_sampler.local_mass().item()))
for k in bad[:5].tolist():
  say("   sample", k, int(sh[k]), int(ref[k]))

# 2. losses: sharded _expectation vs the same ranks computing everything locally (same sharded sampler)
ham = cq.convert_to_tensor([arch.tfim_ring(qubits)])
beta = torch.tensor(0.7, device=dev)
q1, _, _ = make_qhbm("v", 21)
l_sh = float(inference.vqt(q1, ham, beta))
q2, _, _ = make_qhbm("v", 21)
with qd.local_shard():
  l_loc = float(inference.vqt(q2, ham, beta))
say("vqt  sharded", repr(l_sh), "local", repr(l_loc), "rel", abs(l_sh - l_loc) / abs(l_loc))
d1, _, _ = make_qhbm("d", 31)
m1, _, _ = make_qhbm("m", 41)
l_sh = float(inference.qmhl(qdata.QHBMData(d1), m1))
d2, _, _ = make_qhbm("d", 31)
m2, _, _ = make_qhbm("m", 41)
with qd.local_shard():
  l_loc = float(inference.qmhl(qdata.QHBMData(d2), m2))
say("qmhl sharded", repr(l_sh), "local", repr(l_loc), "rel", abs(l_sh - l_loc) / abs(l_loc))
# 3. pieces of qmhl
d3, _, _ = make_qhbm("d", 31)
m3, _, _ = make_qhbm("m", 41)
a = float(qdata.QHBMData(d3).expectation(m3.modular_hamiltonian))
b = float(m3.e_inference.log_partition())
d4, _, _ = make_qhbm("d", 31)
m4, _, _ = make_qhbm("m", 41)
with qd.local_shard():
  a2 = float(qdata.QHBMData(d4).expectation(m4.modular_hamiltonian))
  b2 = float(m4.e_inference.log_partition())
say("qmhl pieces: <H> sharded", repr(a), "local", repr(a2), "| logZ sharded", repr(b), "local", repr(b2))
dist.destroy_process_group()
