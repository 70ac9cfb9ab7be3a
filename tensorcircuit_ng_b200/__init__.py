"""
tensorcircuit_ng_b200 — a B200-native engine behind TensorCircuit-NG's contractor
plugin surface (statevector evolution + sliced tensor-network contraction).

Host code is Python/PyTorch; all arithmetic on states and tensor networks is done by
hand-written sm_100a CUDA reached through the C ABI in include/tcb200.h
(lib/libtcb200.so, loaded with ctypes).  There is no CPU fallback.

    import tensorcircuit_ng_b200 as tc
    c = tc.Circuit(30); c.h(range(30)); ...; c.expectation_ps(z=[0, 1])
"""

__version__ = "0.1.0"

from . import _lib  # noqa: F401
from . import tn  # noqa: F401
from . import gates  # noqa: F401
from . import cons  # noqa: F401
from .cons import (  # noqa: F401
    set_contractor,
    get_contractor,
    runtime_contractor,
    set_function_contractor,
)
from .gates import num_to_tensor, array_to_tensor  # noqa: F401
from . import circuit  # noqa: F401
from .circuit import Circuit, expectation  # noqa: F401
from . import simplify  # noqa: F401
from . import planner  # noqa: F401
from . import passplan  # noqa: F401
from . import svengine  # noqa: F401
from . import tnengine  # noqa: F401
from . import expect  # noqa: F401
from . import autograd  # noqa: F401
from . import sharded  # noqa: F401
from . import experimental  # noqa: F401
from . import backend  # noqa: F401
from . import quantum  # noqa: F401
from . import sampling  # noqa: F401
from . import interfaces  # noqa: F401
from . import torchnn  # noqa: F401
from .torchnn import QuantumNet, TorchLayer  # noqa: F401
from . import templates  # noqa: F401
