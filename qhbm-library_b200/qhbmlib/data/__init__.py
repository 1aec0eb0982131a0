"""qhbmlib.data (mirror of /root/reference/qhbmlib/data/__init__.py)."""
from qhbmlib.data.qhbm_data import QHBMData
from qhbmlib.data.quantum_data import QuantumData

__all__ = ["QHBMData", "QuantumData"]
