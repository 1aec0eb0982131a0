"""cProfile of the VQT training step through the Python API (host-side overhead; GPU work is ~1.9 ms)."""
import cProfile
import os
import pstats
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

n, num_samples = 16, 500
qubits = cq.GridQubit.rect(1, n)
energy = models.KOBE(list(range(n)), 2, energy_utils.RandomNormal(0.0, 0.1, 4))
e_infer = inference.AnalyticEnergyInference(energy, num_samples)
circ = models.DirectQuantumCircuit(arch.get_hardware_efficient_model_unitary(qubits, 2, "q"),
                                   energy_utils.RandomUniform(-1, 1, 11))
qhbm = inference.QHBM(e_infer, inference.AnalyticQuantumInference(circ))
h = cq.convert_to_tensor([arch.tfim_ring(qubits)])
beta = torch.tensor(1.0, device="cuda")
opt = torch.optim.Adam(qhbm.trainable_variables, lr=1e-2)


def step():
  opt.zero_grad()
  loss = inference.vqt(qhbm, h, beta)
  loss.backward()
  opt.step()


for _ in range(10):
  step()
torch.cuda.synchronize()
# host time per step WITHOUT waiting for the GPU (enqueue cost) and with
t0 = time.perf_counter()
for _ in range(50):
  step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"50 steps: host enqueue {1e3 * (t1 - t0) / 50:.2f} ms/step, incl. final drain {1e3 * (t2 - t0) / 50:.2f} ms/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
  step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
st.sort_stats("tottime").print_stats(25)
