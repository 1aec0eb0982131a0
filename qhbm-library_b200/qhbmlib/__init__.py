"""qhbmlib on B200: the reference's Python API over hand-written sm_100a kernels.

Sub-packages mirror /root/reference/qhbmlib (`models`, `inference`, `data`, `utils`);
`engine` and `_native` are the binding to libqhbm_b200.so.
"""
__version__ = "0.3.0+b200.1"
