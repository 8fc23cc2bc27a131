"""Arguments / config of the SCD adapted-sampling parity fixture, shared by tests/golden/make_golden.py
(which feeds them to the reference's factory) and by the tests (which feed them to this package's)."""
import argparse
from types import SimpleNamespace as NS


def adapted_args(dc_type):
    """Arguments of run_adapted_sampling.py (reference :17-37) shrunk to a CPU-sized chain."""
    return argparse.Namespace(
        method='dds', num_steps=6, adapt_freq=2, eta=0.85, gamma=0.05, adaptation='full',
        lora_include_blocks=None, lora_rank=2, tv_penalty=1e-3, num_optim_step=3, lr=1e-3,
        add_cg=True, dc_type=dc_type, cg_iter=2, early_stopping_pct=1.0)


def adapted_config(batch, device='cpu'):
    return NS(device=device, sampling=NS(batch_size=batch, eps=1e-3, travel_length=1, travel_repeat=1),
              model=NS(in_channels=1))
