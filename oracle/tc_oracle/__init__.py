"""
ORACLE — CPU (numpy) restatement of TensorCircuit-NG's contraction hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under `tensorcircuit_ng_b200/` imports this
package; only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs do, and only as the checker / baseline.

Parity status: numerics are pinned by the reference's own
golden values (tests/golden/reference_kats.json, checked in
tests/test_oracle_golden.py).  Contraction-plan bit-exactness against
opt_einsum / cotengra is PARITY UNPINNED: those planners are third-party
packages absent from this image and no reference test fixes a concrete path.
"""
from . import tn, gates, paths, cons, circuit, quantum, treeexec  # noqa: F401
from .circuit import Circuit  # noqa: F401
from .cons import set_contractor, runtime_contractor  # noqa: F401
