"""
ctypes binding of the C-ABI library `lib/libtcb200.so` (include/tcb200.h).

There is no CPU fallback: if the shared library is missing, or a compute entry
point is called without a CUDA device, this module raises.
"""

from __future__ import annotations

import ctypes
import os
from ctypes import (
    POINTER,
    Structure,
    c_char_p,
    c_double,
    c_int,
    c_int8,
    c_int32,
    c_int64,
    c_uint64,
    c_void_p,
)
from typing import Any, Optional, Sequence

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TCB_LIB_PATH") or os.path.join(_HERE, "lib", "libtcb200.so")

_lib: Optional[ctypes.CDLL] = None


class EngineError(RuntimeError):
    pass


class ContractDesc(Structure):
    _fields_ = [
        ("n_batch", c_int32),
        ("n_m", c_int32),
        ("n_n", c_int32),
        ("n_k", c_int32),
        ("batch_a", c_int8 * 32),
        ("batch_b", c_int8 * 32),
        ("batch_c", c_int8 * 32),
        ("m_a", c_int8 * 32),
        ("m_c", c_int8 * 32),
        ("n_b", c_int8 * 32),
        ("n_c", c_int8 * 32),
        ("k_a", c_int8 * 32),
        ("k_b", c_int8 * 32),
        ("conj_a", c_int32),
        ("conj_b", c_int32),
    ]


# name -> (restype, argtypes); every symbol declared in include/tcb200.h
SYMBOLS = {
    "tcb_abi_version": (c_int, []),
    "tcb_last_error": (c_char_p, []),
    "tcb_release_scratch": (c_int, []),
    "tcb_device_info": (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_uint64)]),
    "tcb_sv_init_zero": (c_int, [c_void_p, c_int, c_int64, c_void_p]),
    "tcb_sv_init_product": (c_int, [c_void_p, c_int, c_void_p, c_int, c_uint64, c_void_p]),
    "tcb_sv_apply_dense": (
        c_int,
        [c_void_p, c_int, c_int64, POINTER(c_int), c_int, c_void_p, c_int64, c_void_p],
    ),
    "tcb_sv_apply_diag": (
        c_int,
        [c_void_p, c_int, c_int64, POINTER(c_int), c_int, c_void_p, c_int64, c_int64, c_uint64, c_void_p],
    ),
    "tcb_sv_run_pass": (
        c_int,
        [c_void_p, c_int, c_int64, c_void_p, c_int32, c_int, c_int, c_int, c_void_p, c_int64, c_uint64,
         c_void_p],
    ),
    "tcb_sv_run_pass_generate": (
        c_int,
        [c_void_p, c_int, c_void_p, c_int32, c_int, c_int, c_int, c_void_p, c_uint64, c_void_p, c_int, c_void_p],
    ),
    "tcb_sv_run_pass_oop": (
        c_int,
        [c_void_p, c_void_p, c_int, c_int64, c_void_p, c_int32, c_int, c_int, c_int, c_void_p, c_int64,
         c_uint64, c_void_p],
    ),  # fmt: skip
    "tcb_sv_expect_z": (
        c_int,
        [c_void_p, c_int, c_int64, c_void_p, c_int, c_uint64, c_void_p, c_void_p],
    ),
    "tcb_sv_expect_pauli": (
        c_int,
        [c_void_p, c_int, c_int64, c_uint64, c_uint64, c_int, c_uint64, c_void_p, c_void_p],
    ),
    "tcb_sv_pauli_sum": (
        c_int,
        [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_uint64, c_void_p, c_int, c_void_p,
         c_void_p],
    ),  # fmt: skip
    "tcb_sv_inner": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p]),
    "tcb_sv_gate_grad": (
        c_int,
        [c_void_p, c_void_p, c_int, c_int64, POINTER(c_int), c_int, c_void_p, c_int64, c_void_p],
    ),
    "tcb_sv_adjoint_step": (
        c_int,
        [c_void_p, c_void_p, c_int, c_int64, POINTER(c_int), c_int, c_void_p, c_int64, c_void_p, c_int64, c_void_p],
    ),
    "tcb_sv_cross_marginals": (
        c_int,
        [c_void_p, c_void_p, c_int, c_int64, c_int, POINTER(c_int), c_void_p, c_int64, c_void_p],
    ),
    "tcb_sv_cross_rdm": (
        c_int,
        [c_void_p, c_void_p, c_int, c_int64, c_int, POINTER(c_int), c_int, c_void_p, c_int64, c_void_p],
    ),
    "tcb_sv_sample_prepare": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "tcb_sv_sample": (
        c_int,
        [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p],
    ),
    "tcb_sv_pack_half": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "tcb_sv_unpack_half": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "tcb_sv_pack_bits": (
        c_int,
        [c_void_p, c_void_p, c_int, c_int, POINTER(c_int), c_uint64, c_uint64, c_uint64, c_void_p],
    ),
    "tcb_sv_unpack_bits": (
        c_int,
        [c_void_p, c_void_p, c_int, c_int, POINTER(c_int), c_uint64, c_uint64, c_uint64, c_void_p],
    ),
    "tcb_tn_contract": (
        c_int,
        [c_void_p, c_int64, c_void_p, c_int64, c_void_p, POINTER(ContractDesc), c_int, c_void_p],
    ),
    "tcb_sv_plan_create": (
        c_int,
        [c_int, c_void_p, c_int64, c_void_p, c_int, c_void_p, c_int, POINTER(c_void_p)],
    ),
    "tcb_sv_plan_destroy": (c_int, [c_void_p]),
    "tcb_sv_plan_workspace_size": (c_int64, [c_void_p]),
    "tcb_sv_plan_execute": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_uint64, c_void_p]),
    "tcb_sv_plan_vjp": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "tcb_sv_plan_vjp_range": (
        c_int,
        [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    ),
    "tcb_sv_plan_launches": (c_int, [c_void_p, c_int]),
    "tcb_tn_plan_create": (
        c_int,
        [c_int, c_void_p, c_int, c_void_p, POINTER(ContractDesc), c_void_p, c_int, c_void_p, c_void_p,
         POINTER(c_void_p)],
    ),  # fmt: skip
    "tcb_tn_plan_destroy": (c_int, [c_void_p]),
    "tcb_tn_plan_workspace_size": (c_int64, [c_void_p]),
    "tcb_tn_plan_output_elems": (c_int64, [c_void_p]),
    "tcb_tn_plan_execute": (c_int, [c_void_p, POINTER(c_void_p), c_uint64, c_void_p, c_void_p, c_int64, c_void_p]),
    "tcb_tn_plan_launches": (c_int, [c_void_p]),
}


def load() -> ctypes.CDLL:
    """Load the shared library (no CUDA device needed just to load and list symbols)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  tensorcircuit_ng_b200 has no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.tcb_abi_version() != 1:
        raise EngineError("libtcb200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().tcb_last_error().decode("utf-8", "replace")
        raise EngineError(f"libtcb200: {msg} (rc={rc})")


def require_cuda(t: Any, what: str = "tensor") -> None:
    if not getattr(t, "is_cuda", False):
        raise EngineError(
            f"{what} lives on {getattr(t, 'device', '?')}: the B200 engine only runs on CUDA tensors "
            "(no CPU fallback). Use torch.set_default_device('cuda') or move the inputs."
        )


def stream_ptr() -> int:
    import torch

    return int(torch.cuda.current_stream().cuda_stream)


def int_array(vals: Sequence[int]) -> Any:
    return (c_int * len(vals))(*[int(v) for v in vals])


launch_count = 0  # kernels launched through the C ABI (bench.py reports it as gpu_launches)


def call(name: str, *args: Any) -> None:
    global launch_count
    check(getattr(load(), name)(*args))
    launch_count += 1
