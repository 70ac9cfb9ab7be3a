"""
The slice of the backend surface the hot path's callers use for differentiation
(`tensorcircuit/backends/pytorch_backend.py:775-786,816-878`: `value_and_grad`, `vmap`, `vvag`).

Gradients come from torch autograd through the engine's own `autograd.Function`s (adjoint-mode
statevector vjp: the backward pass re-runs the fused pass kernels with U^dagger; tensor-network
steps differentiate as two more contractions).  `vmap` / `vvag` evaluate the function ONCE under
`torch.vmap` when it stays inside what the batched engine path covers (circuits from |0...0>, Pauli-sum
energies: one circuit build, kernels launched with `batch` = B) and otherwise fall back to a loop over
the vectorised argument — the semantics of the reference's numpy backend
(`tensorcircuit/backends/numpy_backend.py:540-564`) with the reference's torch-backend signature.
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


batched_mode = __import__("os").environ.get("TCB_VMAP", "auto")  # auto | loop | strict
last_vmap_path = ""  # "batched" or "loop: <why>" — what the most recent vmap / vvag call did


def _try_batched(f: Callable[..., Any], args: Sequence[Any], vnums: Tuple[int, ...], kws: Any) -> Any:
    """Evaluate f once under torch.vmap (functorch batches the torch ops on the parameters; the engine runs
    its kernels with batch = B, see autograd.evolve).  Chunked so that the B states fit in free HBM."""
    n = args[vnums[0]].shape[0]
    in_dims = tuple(0 if i in vnums else None for i in range(len(args)))
    chunk = n
    if torch.cuda.is_available():
        free, _ = torch.cuda.mem_get_info()
        chunk = max(1, min(n, _batched_chunk_hint(free)))
    outs = []
    for b0 in range(0, n, chunk):
        part = [x[b0 : b0 + chunk] if i in vnums else x for i, x in enumerate(args)]
        outs.append(torch.vmap(lambda *a: f(*a, **kws), in_dims=in_dims)(*part))
    if isinstance(outs[0], (tuple, list)):
        return tuple(torch.cat([o[k] for o in outs]) for k in range(len(outs[0])))
    return torch.cat(outs) if len(outs) > 1 else outs[0]


state_bytes_hint = 0  # set by callers that know the state size (bytes of ONE state); 0 = assume it fits


def _batched_chunk_hint(free_bytes: int) -> int:
    if state_bytes_hint <= 0:
        return 1 << 30
    return max(1, int(free_bytes * 0.8) // (6 * state_bytes_hint))  # psi, lam, H psi, clones ...


def vmap(f: Callable[..., Any], vectorized_argnums: Union[int, Sequence[int]] = 0) -> Callable[..., Any]:
    """pytorch_backend.py:816-828.  One batched evaluation under torch.vmap when the function stays inside what
    the batched engine path covers; otherwise the loop of the reference's numpy backend
    (numpy_backend.py:540-564)."""
    vnums = _as_tuple(vectorized_argnums)

    def loop(*args: Any, **kws: Any) -> Any:
        n = args[vnums[0]].shape[0]
        outs = []
        for b in range(n):
            a = [x[b] if i in vnums else x for i, x in enumerate(args)]
            outs.append(f(*a, **kws))
        if isinstance(outs[0], (tuple, list)):
            return tuple(torch.stack([o[k] for o in outs]) for k in range(len(outs[0])))
        return torch.stack(outs)

    def wrapper(*args: Any, **kws: Any) -> Any:
        global last_vmap_path
        if batched_mode != "loop" and all(isinstance(args[i], torch.Tensor) for i in vnums):
            try:
                out = _try_batched(f, args, vnums, kws)
                last_vmap_path = "batched"
                return out
            except Exception as e:  # pylint: disable=broad-except  (anything functorch / the engine cannot batch)
                if batched_mode == "strict":
                    raise
                last_vmap_path = f"loop: {type(e).__name__}: {e}"[:300]
        else:
            last_vmap_path = "loop: requested"
        return loop(*args, **kws)

    return wrapper


def vectorized_value_and_grad(f: Callable[..., Any], argnums: Union[int, Sequence[int]] = 0,
                              vectorized_argnums: Union[int, Sequence[int]] = 0, has_aux: bool = False) -> Callable[..., Any]:  # fmt: skip
    """pytorch_backend.py:830-878: values [B], gradient of sum_b f(x_b) w.r.t. args[argnums] — stacked
    per-sample gradients when the differentiated argument is itself vectorised, summed otherwise.
    Fast path: ONE evaluation under torch.vmap and one backward of sum_b f(x_b) (samples are independent, so
    that is exactly the stacked per-sample gradient); fallback: the loop."""
    nums, vnums = _as_tuple(argnums), _as_tuple(vectorized_argnums)
    vag = value_and_grad(f, argnums=nums, has_aux=has_aux)

    def loop(*args: Any, **kws: Any) -> Any:
        n = args[vnums[0]].shape[0]
        vals, grads = [], None
        for b in range(n):
            a = [x[b] if i in vnums else x for i, x in enumerate(args)]
            v, g = vag(*a, **kws)
            vals.append(v[0] if has_aux else v)
            if grads is None:
                grads = [[] for _ in nums]
            for k, gk in enumerate(g):  # (`vag` was built with the tuple `nums`: always a tuple)
                grads[k].append(gk)
        out = []
        for k, i in enumerate(nums):
            st = torch.stack(grads[k])
            out.append(st if i in vnums else st.sum(0))
        g = out[0] if isinstance(argnums, int) else tuple(out)
        return torch.stack(vals), g

    def batched(*args: Any, **kws: Any) -> Any:
        args = list(args)
        xs = []
        for i in nums:
            x = args[i] if isinstance(args[i], torch.Tensor) else torch.as_tensor(args[i])
            x = x.detach().clone().requires_grad_(True)
            args[i] = x
            xs.append(x)
        vals = _try_batched(f, args, vnums, kws)
        if isinstance(vals, (tuple, list)):
            raise TypeError("tuple outputs take the loop")
        if vals.is_complex():
            vals = vals.real
        gs = torch.autograd.grad(vals.sum(), xs, allow_unused=True)
        gs = tuple(torch.zeros_like(x) if g is None else g for g, x in zip(gs, xs))
        return vals.detach(), (gs[0] if isinstance(argnums, int) else gs)

    def wrapper(*args: Any, **kws: Any) -> Any:
        global last_vmap_path
        ok = batched_mode != "loop" and not has_aux and all(isinstance(args[i], torch.Tensor) for i in vnums)
        if ok:
            try:
                out = batched(*args, **kws)
                last_vmap_path = "batched"
                return out
            except Exception as e:  # pylint: disable=broad-except
                if batched_mode == "strict":
                    raise
                last_vmap_path = f"loop: {type(e).__name__}: {e}"[:300]
        else:
            last_vmap_path = "loop: requested"
        return loop(*args, **kws)

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
