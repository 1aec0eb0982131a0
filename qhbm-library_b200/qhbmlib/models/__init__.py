"""qhbmlib.models (mirror of /root/reference/qhbmlib/models/__init__.py)."""
from qhbmlib.models.circuit import DirectQuantumCircuit
from qhbmlib.models.circuit import QAIA
from qhbmlib.models.circuit import QuantumCircuit
from qhbmlib.models.energy import BernoulliEnergy
from qhbmlib.models.energy import BitstringEnergy
from qhbmlib.models.energy import KOBE
from qhbmlib.models.energy import PauliMixin
from qhbmlib.models.energy_utils import Parity
from qhbmlib.models.energy_utils import SpinsFromBitstrings
from qhbmlib.models.energy_utils import VariableDot
from qhbmlib.models.hamiltonian import Hamiltonian

__all__ = ["BernoulliEnergy", "BitstringEnergy", "DirectQuantumCircuit", "Hamiltonian", "KOBE", "Parity",
           "PauliMixin", "QAIA", "QuantumCircuit", "SpinsFromBitstrings", "VariableDot"]
