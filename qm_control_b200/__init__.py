"""qm_control_b200 — B200-native batched MPC + whole-body-control solver (drop-in for qm_control's hot path).

The product is the CUDA library ``libqmb200.so`` behind the C ABI in ``include/qmb200.h``; this package is the
Python mirror of the reference's interfaces for that path (QMInterface, WbcBase/HierarchicalWbc, the SQP MPC seam)
used by tests and bench.py.  There is no CPU fallback: importing works without a GPU (so the build and the symbol
checks run anywhere), but constructing a solver without a CUDA device raises.
"""
from ._lib import LIB_PATH, ASSETS, load_library, QmbError  # noqa: F401
from .interface import QMInterface, Solver  # noqa: F401
from .wbc import HierarchicalWbc, HierarchicalMpcWbc  # noqa: F401
from .mpc import SqpMpc  # noqa: F401
from .controller import QMController, QMMpcController  # noqa: F401

__all__ = ["QMInterface", "Solver", "HierarchicalWbc", "HierarchicalMpcWbc", "SqpMpc", "QMController", "QMMpcController", "load_library", "QmbError", "LIB_PATH", "ASSETS"]
