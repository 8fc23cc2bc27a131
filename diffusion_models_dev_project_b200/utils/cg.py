"""Batched conjugate gradient for ``(I + gamma A*A) x = rhs``.

Signature of the reference's ``cg(op, x, rhs, n_iter, tol)``
(src/utils/cg.py:11-39): fixed ``n_iter`` iterations from the start value ``x``,
per-sample step sizes over dims [1,2,3], no tolerance test (``tol`` is accepted
and, as in the reference, unused) and no guard against a zero residual.

When ``op`` is a :class:`~..physics.b200_ray_trafo.NormalOp` on a
:class:`B200RayTrafo` and no gradient is required, the whole solve runs as the
fused CUDA launch sequence ``scd_cg`` (A, A* with the axpy and <p,d> in its
epilogue, one update kernel, one direction kernel per iteration).  Any other
callable -- and every call that must be differentiated, e.g. the LoRA
adaptation of SCD (reference src/samplers/utils.py:241-260) -- runs the same
recurrences as tensor operations on whatever device the tensors live on.
"""
import torch
from torch import Tensor


def _fused_ok(op, x: Tensor, rhs: Tensor) -> bool:
    from ..physics.b200_ray_trafo import B200RayTrafo, NormalOp
    if not isinstance(op, NormalOp) or not isinstance(op.ray_trafo, B200RayTrafo):
        return False
    if torch.is_grad_enabled() and (x.requires_grad or rhs.requires_grad):
        return False
    return x.is_cuda and x.dtype == torch.float32 and x.dim() == 4


def cg(op: callable, x: Tensor, rhs: Tensor, n_iter: int = 5, tol: float = 1e-10) -> Tensor:
    if _fused_ok(op, x, rhs):
        return op.ray_trafo.cg_solve(x, rhs, op.gamma, n_iter)

    def bdot(a, b):
        return (a * b).sum(dim=[1, 2, 3])

    r = rhs - op(x)
    p = r
    rr = bdot(r, r)
    for _ in range(n_iter):
        d = op(p)
        alpha = rr / bdot(p, d)
        x = x + alpha[:, None, None, None] * p
        r = r - alpha[:, None, None, None] * d
        rr_new = bdot(r, r)
        beta = rr_new / rr
        rr = rr_new
        p = r + beta[:, None, None, None] * p
    return x
