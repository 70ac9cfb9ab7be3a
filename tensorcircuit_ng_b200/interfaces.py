"""
`tensorcircuit/interfaces/torch.py:17-125` for an engine whose tensors already ARE torch tensors.

The reference wraps a quantum function of another ML backend in a `torch.autograd.Function` (dlpack in,
vjp out).  Here the function runs on torch directly and autograd flows through the engine's own
`autograd.Function`s (adjoint statevector walk, Pauli-sum / contraction vjps), so the interface only
places host tensors on the GPU and returns results on the inputs' device.
"""

from __future__ import annotations

from typing import Any, Callable

import torch


def _to(x: Any, device: torch.device) -> Any:
    if isinstance(x, torch.Tensor):
        return x.to(device)
    if isinstance(x, (list, tuple)):
        return type(x)(_to(v, device) for v in x)
    if isinstance(x, dict):
        return {k: _to(v, device) for k, v in x.items()}
    return x


def _first_device(x: Any) -> Any:
    if isinstance(x, torch.Tensor):
        return x.device
    if isinstance(x, (list, tuple)):
        for v in x:
            d = _first_device(v)
            if d is not None:
                return d
    if isinstance(x, dict):
        return _first_device(list(x.values()))
    return None


def torch_interface(fun: Callable[..., Any], jit: bool = False, enable_dlpack: bool = False) -> Callable[..., Any]:
    """Same call signature as the reference; `jit` / `enable_dlpack` have nothing to do here."""

    def wrapped(*x: Any) -> Any:
        src = _first_device(x)
        if not torch.cuda.is_available():
            from . import _lib

            raise _lib.EngineError("torch_interface: no CUDA device (the B200 engine has no CPU fallback)")
        dev = src if src is not None and src.type == "cuda" else torch.device("cuda", torch.cuda.current_device())
        y = fun(*_to(x, dev))
        return _to(y, src) if src is not None and src != dev else y

    return wrapped


pytorch_interface = torch_interface
