"""Minimal LoRA injector for the benchmark / example score models (caller-side PyTorch code).

The SCD predictors only look for modules whose class is named ``LoraInjectedLinear`` /
``LoraInjectedConv2d`` and toggle their ``scale`` attribute (reference src/samplers/utils.py:262-278);
the reference takes those classes from its vendored ``lora`` package, which stays outside this repo.
This stand-in wraps every ``nn.Linear`` and 1x1 / 3x3 ``nn.Conv2d`` found under the named blocks with a
rank-``r`` residual branch ``scale * up(down(x))`` (``up`` zero-initialised: injection leaves the
model's output unchanged)."""
import torch
from torch import nn


class LoraInjectedLinear(nn.Module):
    def __init__(self, base: nn.Linear, r: int):
        super().__init__()
        self.linear = base
        self.lora_down = nn.Linear(base.in_features, r, bias=False)
        self.lora_up = nn.Linear(r, base.out_features, bias=False)
        nn.init.normal_(self.lora_down.weight, std=1.0 / r)
        nn.init.zeros_(self.lora_up.weight)
        self.scale = 1.0

    def forward(self, x):
        return self.linear(x) + self.scale * self.lora_up(self.lora_down(x))


class LoraInjectedConv2d(nn.Module):
    def __init__(self, base: nn.Conv2d, r: int):
        super().__init__()
        self.conv = base
        self.lora_down = nn.Conv2d(base.in_channels, r, base.kernel_size, base.stride, base.padding, bias=False)
        self.lora_up = nn.Conv2d(r, base.out_channels, 1, bias=False)
        nn.init.normal_(self.lora_down.weight, std=1.0 / r)
        nn.init.zeros_(self.lora_up.weight)
        self.scale = 1.0

    def forward(self, x):
        return self.conv(x) + self.scale * self.lora_up(self.lora_down(x))


def inject_trainable_lora(model: nn.Module, r: int = 4, include_blocks=None, **_):
    """Wrap the Linear / Conv2d layers (below the attributes named in ``include_blocks`` if given) and
    make only the LoRA branches trainable.  Returns the list of new parameters."""
    roots = [model] if not include_blocks else [getattr(model, b) for b in include_blocks if hasattr(model, b)]
    roots = roots or [model]
    new_params = []
    for root in roots:
        for parent in list(root.modules()):
            for name, child in list(parent.named_children()):
                if isinstance(child, nn.Linear):
                    wrapped = LoraInjectedLinear(child, r)
                elif isinstance(child, nn.Conv2d) and child.groups == 1:
                    wrapped = LoraInjectedConv2d(child, r)
                else:
                    continue
                wrapped.to(next(child.parameters()).device)
                setattr(parent, name, wrapped)
                for p in (*wrapped.lora_down.parameters(), *wrapped.lora_up.parameters()):
                    p.requires_grad_(True)
                    new_params.append(p)
    return new_params
