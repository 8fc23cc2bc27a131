"""Small end-to-end pass over every kernel of libscd_b200.so, meant to be run under compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_run.py
    compute-sanitizer --tool racecheck python tools/sanitize_run.py
    compute-sanitizer --tool synccheck python tools/sanitize_run.py

(compute-sanitizer is closed on the GPU pool this repository is developed on -- the out-of-bounds write check
there is tests/test_gpu_guards.py; this script stays as the command to use where the tool is available.)
Sizes are small (the tools slow kernels down 10-100x) but cover every sample-group width, the cluster row
split, ragged groups, the interleaved-sinogram pair, CG, the fused DDS step and the loss kernels.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffusion_models_dev_project_b200 as pkg  # noqa: E402
from diffusion_models_dev_project_b200 import fused  # noqa: E402


def main():
    dev = torch.device('cuda')
    torch.manual_seed(0)
    for shape, na in (((64, 64), 12), ((33, 47), 5), ((256, 256), 8)):
        rt = pkg.B200RayTrafo(shape, na)
        for b in (1, 2, 3, 5, 9, 17):
            if shape[0] == 256 and b not in (1, 9):
                continue
            x = torch.rand(b, 1, *shape, device=dev)
            y = rt(x)
            z = rt.trafo_adjoint(y)
            op = rt.normal_op(0.05)
            w = pkg.cg(op, x, z, 2)
            assert torch.isfinite(y).all() and torch.isfinite(z).all() and torch.isfinite(w).all()
        sde = pkg.DDPM()
        b = 3
        x = torch.randn(b, 1, *shape, device=dev)
        s = torch.randn(b, 1, *shape, device=dev)
        t = torch.ones(b, device=dev) * 500.
        tp = torch.ones(b, device=dev) * 490.
        atb = rt.trafo_adjoint(rt(torch.rand(b, 1, *shape, device=dev)))
        out = pkg.decomposed_diffusion_sampling_sde_predictor(
            score=lambda v, tt: s, sde=sde, x=x, time_step=(t, tp), eta=0.15, gamma=0.05, step_size=1,
            use_simplified_eqn=True, ray_trafo=rt, rhs=atb, cg_kwargs={'max_iter': 2})
        assert torch.isfinite(out[0]).all()
        abar = sde.alpha_bar_table(dev)
        fused.tweedie_rhs(x, s, t, abar, atb=atb, gamma=0.05)
        fused.ddim_ddpm(x, s, atb, t, tp, abar, 0.15)
        xr = torch.rand(b, 1, *shape, device=dev, requires_grad=True)
        loss = pkg.adaptation_loss(xr, rt(x).detach(), rt, 1e-3)
        loss.backward()
        assert torch.isfinite(xr.grad).all()
        f = rt.fbp(rt(x))
        assert torch.isfinite(f).all()
    torch.cuda.synchronize()
    print('sanitize_run: ok')


if __name__ == '__main__':
    main()
