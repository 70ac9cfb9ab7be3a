"""
The slice of the backend surface the hot path's callers use for differentiation
(`tensorcircuit/backends/pytorch_backend.py:775-786,816-878`: `value_and_grad`, `vmap`, `vvag`).

Gradients come from torch autograd through the engine's own `autograd.Function`s (adjoint-mode
statevector vjp: the backward pass re-runs the fused pass kernels with U^dagger; tensor-network
steps differentiate as two more contractions).  `vmap` / `vvag` evaluate the batch as a loop over
the vectorised argument — the semantics of the reference's numpy backend
(`tensorcircuit/backends/numpy_backend.py:540-564`) with the reference's torch-backend signature;
a functorch batching rule for the pass kernel (one launch with `batch` > 1, which the C ABI already
supports) is the follow-up.
"""

from __future__ import annotations

from typing import Any, Callable, Sequence, Tuple, Union

import torch


def _as_tuple(x: Union[int, Sequence[int]]) -> Tuple[int, ...]:
    return (x,) if isinstance(x, int) else tuple(x)


def value_and_grad(f: Callable[..., Any], argnums: Union[int, Sequence[int]] = 0, has_aux: bool = False) -> Callable[..., Any]:
    """pytorch_backend.py:775-786.  Returns (value, grad) with grad shaped like args[argnums]."""
    nums = _as_tuple(argnums)

    def wrapper(*args: Any, **kws: Any) -> Any:
        args = list(args)
        xs = []
        for i in nums:
            x = args[i]
            x = x if isinstance(x, torch.Tensor) else torch.as_tensor(x)
            x = x.detach().clone().requires_grad_(True)
            args[i] = x
            xs.append(x)
        out = f(*args, **kws)
        v = out[0] if has_aux else out
        if v.is_complex():
            v = v.real  # the reference differentiates the real part of a complex scalar on torch
        gs = torch.autograd.grad(v, xs, allow_unused=True)
        gs = tuple(torch.zeros_like(x) if g is None else g for g, x in zip(gs, xs))
        g = gs[0] if isinstance(argnums, int) else gs
        if has_aux:
            return (out[0].detach(), *out[1:]), g
        return v.detach(), g

    return wrapper


def vmap(f: Callable[..., Any], vectorized_argnums: Union[int, Sequence[int]] = 0) -> Callable[..., Any]:
    """pytorch_backend.py:816-828 (loop semantics, numpy_backend.py:540-564)."""
    vnums = _as_tuple(vectorized_argnums)

    def wrapper(*args: Any, **kws: Any) -> Any:
        n = args[vnums[0]].shape[0]
        outs = []
        for b in range(n):
            a = [x[b] if i in vnums else x for i, x in enumerate(args)]
            outs.append(f(*a, **kws))
        if isinstance(outs[0], (tuple, list)):
            return tuple(torch.stack([o[k] for o in outs]) for k in range(len(outs[0])))
        return torch.stack(outs)

    return wrapper


def vectorized_value_and_grad(f: Callable[..., Any], argnums: Union[int, Sequence[int]] = 0,
                              vectorized_argnums: Union[int, Sequence[int]] = 0, has_aux: bool = False) -> Callable[..., Any]:  # fmt: skip
    """pytorch_backend.py:830-878: values [B], gradient of sum_b f(x_b) w.r.t. args[argnums] — stacked
    per-sample gradients when the differentiated argument is itself vectorised, summed otherwise."""
    nums, vnums = _as_tuple(argnums), _as_tuple(vectorized_argnums)
    vag = value_and_grad(f, argnums=nums, has_aux=has_aux)

    def wrapper(*args: Any, **kws: Any) -> Any:
        n = args[vnums[0]].shape[0]
        vals, grads = [], None
        for b in range(n):
            a = [x[b] if i in vnums else x for i, x in enumerate(args)]
            v, g = vag(*a, **kws)
            vals.append(v[0] if has_aux else v)
            if grads is None:
                grads = [[] for _ in nums]
            for k, gk in enumerate(g):
                grads[k].append(gk)
        out = []
        for k, i in enumerate(nums):
            st = torch.stack(grads[k])
            out.append(st if i in vnums else st.sum(0))
        g = out[0] if isinstance(argnums, int) else tuple(out)
        return torch.stack(vals), g

    return wrapper


vvag = vectorized_value_and_grad


# -- the handful of array helpers callers of the path write their losses with (pytorch_backend.py) --------
name = "pytorch"
stack = torch.stack
real = torch.real
imag = torch.imag
sum = torch.sum  # noqa: A001  (the reference's backend.sum)
mean = torch.mean
reshape = torch.reshape
ones = torch.ones
zeros = torch.zeros
cast = lambda a, dtype: a.to(getattr(torch, dtype) if isinstance(dtype, str) else dtype)  # noqa: E731


def convert_to_tensor(a: Any) -> torch.Tensor:
    return a if isinstance(a, torch.Tensor) else torch.as_tensor(a)


def numpy(a: Any) -> Any:  # noqa: A001
    return a.detach().resolve_conj().cpu().numpy() if isinstance(a, torch.Tensor) else a


def get_random_state(seed: Any = None) -> torch.Generator:
    """pytorch_backend.py `get_random_state`: a torch.Generator (host side; draws are moved to the GPU)."""
    g = torch.Generator(device="cpu")
    if seed is not None:
        g.manual_seed(int(seed))
    return g


def jit(f: Callable[..., Any], **kws: Any) -> Callable[..., Any]:
    """No tracing compiler on this engine: plans are cached by circuit structure instead (svengine)."""
    return f
