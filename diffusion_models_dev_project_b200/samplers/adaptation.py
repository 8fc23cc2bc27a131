"""Adaptation helpers of SCD (reference src/samplers/adaptation.py)."""
from typing import Dict, Optional

import torch
from torch import nn


def tv_loss(x):
    """Anisotropic total variation on the cropped differences (reference :7-11)."""
    dh = torch.abs(x[..., :, 1:] - x[..., :, :-1])
    dw = torch.abs(x[..., 1:, :] - x[..., :-1, :])
    return torch.sum(dh[..., :-1, :] + dw[..., :, :-1])


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
