"""Config 4 block of bench.py on its own (all ranks):
    python -m torch.distributed.run ... tools/stack_block.py [repeats] [chunk sizes, comma separated]"""
import json, os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench_support import config_blocks as CB      # noqa: E402

world = int(os.environ.get('WORLD_SIZE', '1')); rank = int(os.environ.get('RANK', '0')); local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
torch.set_grad_enabled(False)
chunks = [int(c) for c in sys.argv[2].split(',')] if len(sys.argv) > 2 else [128]
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    for chunk in chunks:
        out = CB.stack_block(dev, world, rank, chunk=chunk)
        if rank == 0:
            print(json.dumps(out))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
