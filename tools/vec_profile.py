"""The HBM-bound vector kernels at a large batch, for an ncu capture:

    ncu --set full --clock-control none -k regex:'tweedie|ddim|cg_update' -s 6 -c 6 python tools/vec_profile.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffusion_models_dev_project_b200 as pkg  # noqa: E402
from diffusion_models_dev_project_b200 import fused  # noqa: E402


def main(batch=256, n=256, angles=60):
    dev = torch.device('cuda')
    torch.set_grad_enabled(False)
    rt = pkg.B200RayTrafo((n, n), angles)
    x = torch.rand(batch, 1, n, n, device=dev)
    s = torch.randn_like(x)
    p = torch.rand_like(x)
    abar = pkg.DDPM().alpha_bar_table(dev)
    t = torch.ones(batch, device=dev) * 500.
    tp = torch.ones(batch, device=dev) * 490.
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    op = rt.normal_op(0.01)
    for _ in range(4):                       # launches 0..5 of each kind warm up, the next ones are captured
        flush.zero_()
        fused.tweedie_rhs(x, s, t, abar, atb=p, gamma=0.01)
        flush.zero_()
        fused.ddim_ddpm(x, s, p, t, tp, abar, 0.15)
        flush.zero_()
        pkg.cg(op, x, p, 1)                  # one cg_update_xr per solve
    torch.cuda.synchronize()


if __name__ == '__main__':
    main()
