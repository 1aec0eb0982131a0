"""Time one VQT training step (loss + backward) through the reference-shaped Python API."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "qhbm-library_b200")):
  sys.path.insert(0, p)
import torch
from qhbmlib import architectures as arch
from qhbmlib import circuits as cq
from qhbmlib import inference
from qhbmlib import models
from qhbmlib.models import energy_utils

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
num_samples = int(sys.argv[2]) if len(sys.argv) > 2 else 500
qubits = cq.GridQubit.rect(1, n)
energy = models.KOBE(list(range(n)), 2, energy_utils.RandomNormal(0.0, 0.1, 4))
e_infer = inference.AnalyticEnergyInference(energy, num_samples)
circ = models.DirectQuantumCircuit(arch.get_hardware_efficient_model_unitary(qubits, 2, "q"),
                                   energy_utils.RandomUniform(-1, 1, 11))
qhbm = inference.QHBM(e_infer, inference.AnalyticQuantumInference(circ))
h = cq.convert_to_tensor([arch.tfim_ring(qubits)])
beta = torch.tensor(1.0, device="cuda")
params = qhbm.trainable_variables
opt = torch.optim.Adam(params, lr=1e-2)
times = []
for it in range(30):
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  opt.zero_grad()
  loss = inference.vqt(qhbm, h, beta)
  loss.backward()
  opt.step()
  torch.cuda.synchronize()
  times.append(time.perf_counter() - t0)
times = sorted(times[10:])
print(f"VQT step (sample, unique, expectation, backward, Adam), n={n}, {num_samples} samples: "
      f"median {times[len(times) // 2] * 1e3:.2f} ms, min {times[0] * 1e3:.2f} ms, last loss {float(loss.detach()):.5f}",
      flush=True)
