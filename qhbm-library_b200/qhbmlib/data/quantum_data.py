"""Interface to quantum data sources (mirror of reference data/quantum_data.py)."""
import abc


class QuantumData(abc.ABC):

  @abc.abstractmethod
  def expectation(self, observable):
    """Scalar expectation of `observable` (OperatorTensor of one PauliSum, or Hamiltonian)."""
    raise NotImplementedError()
