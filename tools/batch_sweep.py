"""BASELINE config 3: batch sweep 8 -> 256 of the data-consistency step (no score model), one GPU or
sample-sharded over N GPUs with no collective.

    python tools/batch_sweep.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29531 tools/batch_sweep.py                  # global batch split evenly over N ranks

Per global batch: device time of one DDS data-consistency step (Tweedie + CG(5) + DDIM = 24 launches, CUDA
events, max over ranks), hot-path samples/s at 100 reverse steps per sample, algorithmic GB/s.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffusion_models_dev_project_b200 as pkg  # noqa: E402

IM, ANGLES, NDET, K = 256, 60, 365, 5
STEP_BYTES = (K + 1) * (2 * 4 * (IM * IM + ANGLES * NDET) + 4 * IM * IM) + K * 6 * 4 * IM * IM + (K - 1) * 3 * 4 * IM * IM \
    + 9 * 4 * IM * IM


def main():
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    torch.set_grad_enabled(False)
    rt = pkg.B200RayTrafo((IM, IM), ANGLES)
    abar = pkg.DDPM().alpha_bar_table(dev)
    rows = []
    for gb in (8, 16, 32, 64, 128, 256):
        if gb % world:
            continue
        B = gb // world
        gen = torch.Generator(device=dev).manual_seed(rank)
        x = torch.rand(B, 1, IM, IM, device=dev, generator=gen)
        s = torch.randn(B, 1, IM, IM, device=dev, generator=gen)
        eps = torch.randn(B, 1, IM, IM, device=dev, generator=gen)
        atb = rt.trafo_adjoint(rt(x))
        t = torch.ones(B, device=dev) * 500.
        tp = torch.ones(B, device=dev) * 490.
        f = lambda: rt.dds_step(x, s, atb, eps, t, tp, abar, 0.01, 0.15, K)          # noqa: E731
        for _ in range(5):
            f()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 30
        e0.record()
        for _ in range(n):
            f()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = float(ms.item())
        rows.append({'global_batch': gb, 'n_gpus': world, 'batch_per_gpu': B, 'ms_per_step': ms,
                     'hot_path_samples_per_s': gb / (100 * ms * 1e-3),
                     'algorithmic_GBps_per_gpu': STEP_BYTES * B / (ms * 1e-3) / 1e9})
    if rank == 0:
        for r in rows:
            print(json.dumps(r))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
