"""Adaptation helpers of SCD (reference src/samplers/adaptation.py)."""
from typing import Dict, Optional

import torch
from torch import nn


def tv_loss(x):
    """Anisotropic total variation on the cropped differences (reference :7-11)."""
    dh = torch.abs(x[..., :, 1:] - x[..., :, :-1])
    dw = torch.abs(x[..., 1:, :] - x[..., :-1, :])
    return torch.sum(dh[..., :-1, :] + dw[..., :, :-1])


class _AdaptationLossFn(torch.autograd.Function):
    """``mean((A x - y)^2) + lam * tv_loss(x)`` with hand-written forward and backward:
    A and the residual/TV reductions forward (5 launches), ``d tv/dx`` and one backprojection
    with the TV gradient as fused addend backward (3 launches).  The gradient through A follows the
    ODL ``OperatorFunction`` pairing used everywhere else: ``(2/N) * A*(r) / c_w``."""

    @staticmethod
    def forward(ctx, x, y, rt, lam):
        import ctypes as C
        from .. import _lib
        lib = _lib.load()
        x = rt._prep(x, rt.im_shape, 'adaptation loss')
        dev = x.device
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        ax = rt._fp(x)
        y = y.to(device=dev, dtype=torch.float32).expand_as(ax).contiguous()
        numel = ax.numel()
        r = torch.empty_like(ax)
        part = torch.empty(int(lib.scd_residual_sq_blocks(numel)), dtype=torch.float32, device=dev)
        images = x.numel() // (rt.im_shape[0] * rt.im_shape[1])
        tvp = torch.empty(images * int(lib.scd_tv_blocks(*rt.im_shape)), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.scd_residual_sq(ax.data_ptr(), y.data_ptr(), r.data_ptr(), part.data_ptr(), numel, stream),
                       'scd_residual_sq')
            _lib.check(lib.scd_tv_loss(x.data_ptr(), tvp.data_ptr(), images, rt.im_shape[0], rt.im_shape[1], stream),
                       'scd_tv_loss')
        ctx.save_for_backward(x, r)
        ctx.rt, ctx.lam, ctx.numel, ctx.images = rt, float(lam), numel, images
        return part.sum() / numel + float(lam) * tvp.sum()

    @staticmethod
    def backward(ctx, g):
        import ctypes as C
        from .. import _lib
        lib = _lib.load()
        x, r = ctx.saved_tensors
        rt = ctx.rt
        dev = x.device
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        tvg = torch.empty_like(x)
        with torch.cuda.device(dev):
            _lib.check(lib.scd_tv_grad(x.data_ptr(), tvg.data_ptr(), ctx.images, rt.im_shape[0], rt.im_shape[1], stream),
                       'scd_tv_grad')
        scale = 2.0 / ctx.numel * rt.adj_scale / rt.geometry.range_weight
        grad = rt._bp(r, scale, addend=tvg, addend_scale=ctx.lam)
        return grad * g, None, None, None


def adaptation_loss(x, observation, ray_trafo, tv_penalty: float):
    """The SCD adaptation loss (reference src/utils/exp_utils.py:256-257).  With a
    :class:`B200RayTrafo` and CUDA tensors it runs as fused kernels with a hand-written
    backward; any other ray transform evaluates the reference's tensor expression."""
    from ..physics.b200_ray_trafo import B200RayTrafo
    if isinstance(ray_trafo, B200RayTrafo) and x.is_cuda and x.dtype == torch.float32:
        return _AdaptationLossFn.apply(x, observation, ray_trafo, float(tv_penalty))
    return torch.mean((ray_trafo(x) - observation).pow(2)) + float(tv_penalty) * tv_loss(x)


def _score_model_adpt(score: nn.Module, impl: str = 'full', adpt_kwargs: Optional[Dict] = None,
                      verbose: bool = True, inject_fn=None) -> None:
    """Select the trainable parameters of the score model (reference :14-52).

    ``impl='lora'`` needs the LoRA injector of the caller's model package (the
    guided-diffusion UNet and its LoRA wrappers stay PyTorch code outside this
    package): pass it as ``inject_fn`` or run inside the reference checkout, where
    ``src.third_party_models.inject_trainable_lora_extended`` is importable."""
    score.requires_grad_(False)
    if impl == 'full':
        score.requires_grad_(True)
    elif impl == 'decoder':
        for part in (score.out, score.output_blocks):
            for name, param in part.named_parameters():
                if "emb_layers" not in name:
                    param.requires_grad = True
    elif impl == 'lora':
        for name, param in score.named_parameters():
            if "bias" in name and "emb_layers" not in name:
                param.requires_grad = True
        if inject_fn is None:
            try:
                from src.third_party_models import inject_trainable_lora_extended as inject_fn
            except ImportError as e:
                raise RuntimeError(
                    "impl='lora' needs a LoRA injector: pass inject_fn=... "
                    "(e.g. the reference's inject_trainable_lora_extended)") from e
        inject_fn(score, **(adpt_kwargs or {}))
    else:
        raise NotImplementedError(impl)

    if verbose:
        num_params = sum(p.numel() for p in score.parameters())
        trainable = sum(p.numel() for p in score.parameters() if p.requires_grad)
        print(f'% of trainable params: {trainable / num_params * 100}')
