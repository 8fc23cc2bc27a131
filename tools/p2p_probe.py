"""Probe: can this box map peer GPU memory into each rank (torch symmetric memory), and what does a
peer write cost?  torchrun --nproc-per-node 2 tools/p2p_probe.py"""
import os
import sys
import time

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    import torch.distributed._symmetric_memory as symm_mem
    n = 64 * 1024 * 1024
    t = symm_mem.empty(n, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(t, group=dist.group.WORLD)
    print(rank, 'rendezvous ok', [hex(p) for p in hdl.buffer_ptrs][:4], 'multicast', hex(hdl.multicast_ptr) if hdl.multicast_ptr else None, flush=True)
    peer = (rank + 1) % world
    pt = hdl.get_buffer(peer, (n,), torch.float32)
    src = torch.full((n,), float(rank + 1), device=dev)
    t.zero_()
    dist.barrier()
    torch.cuda.synchronize()
    pt.copy_(src)                       # P2P write into the peer's buffer
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    exp = float((rank - 1) % world + 1)
    print(rank, 'peer write landed:', float(t[0]), float(t[-1]), 'expected', exp, flush=True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        pt.copy_(src)
    torch.cuda.synchronize()
    dist.barrier()
    a.record()
    for _ in range(10):
        pt.copy_(src)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(rank, 'peer copy %.3f ms for %d MB = %.0f GB/s' % (ms, n * 4 >> 20, n * 4 / ms / 1e6), flush=True)
    # tiny NCCL all-reduce as a cross-rank barrier on the stream
    one = torch.zeros(1, device=dev)
    for _ in range(5):
        dist.all_reduce(one)
    torch.cuda.synchronize()
    a.record()
    for _ in range(20):
        dist.all_reduce(one)
    b.record()
    torch.cuda.synchronize()
    print(rank, 'tiny all_reduce %.1f us' % (a.elapsed_time(b) / 20 * 1e3), flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
