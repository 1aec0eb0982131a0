"""qhbmlib on B200: the reference's Python API over hand-written sm_100a kernels.

Sub-packages mirror /root/reference/qhbmlib (`models`, `inference`, `data`, `utils`);
`circuits` replaces the cirq/TFQ construction surface, `engine` and `_native` bind
libqhbm_b200.so (C ABI: include/qhbm_b200.h).
"""
__version__ = "0.3.0+b200.1"

from qhbmlib import circuits
from qhbmlib import utils
from qhbmlib import models
from qhbmlib import inference
from qhbmlib import data
from qhbmlib import architectures
