"""Multi-GPU partitioning of the hot path (one process per GPU, torch.distributed).

Two cases (the reference itself is single-process, single-GPU):

* **Sample / slice batches** -- every sample's A, A*, CG and DDIM is independent (the CG
  reductions are per sample, reference src/utils/cg.py:22,27,33), so the batch is split into
  contiguous shards and no collective touches the data path: :func:`shard_range`.

* **One large slice stack, angle-sharded** -- rank r owns the angles ``[lo_r, hi_r)``.
  ``A``: each rank computes only its sinogram rows (no communication; the other rows of the
  returned tensor are zero).  ``A*``: each rank backprojects its rows into a full-size partial
  image stack and the partials are summed with an all-reduce (NCCL over NVLink on the GPU box).
  The stack is processed in slice chunks so that the all-reduce of chunk c overlaps the
  backprojection of chunk c+1.  CG vectors are replicated, so every dot product is local.

The wrapped operator only has to provide ``_fp(x, angle_range=)``, ``_bp(y, scale,
angle_range=)``, ``adj_scale``, ``im_shape`` and ``obs_shape`` -- :class:`B200RayTrafo` does;
the CPU tests drive the same code with an oracle-backed stand-in over ``gloo``.
"""
import torch
import torch.distributed as dist
from torch import Tensor


def shard_range(n: int, rank: int, world: int):
    """Contiguous ``[lo, hi)`` share of ``n`` items for ``rank``; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class AngleShardedRayTrafo:
    """Angle-sharded view of a ray transform.

    ``trafo``            -> this rank's rows of ``A x`` (others zero) -- stays sharded.
    ``trafo_adjoint``    -> ``A* y`` summed over ranks (y: full-shape sinogram whose rows outside
                            this rank's range are ignored).
    ``normal_apply``     -> ``v + gamma * A*(A v)`` with one all-reduce per call.
    ``normal_op(gamma)`` -> callable for :func:`..utils.cg.cg` (tensor-op recurrences on replicated vectors).
    """

    def __init__(self, base, group=None, chunk: int = 64):
        self.base = base
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.im_shape = base.im_shape
        self.obs_shape = base.obs_shape
        self.angle_range = shard_range(base.obs_shape[0], self.rank, self.world)
        self.chunk = int(chunk)
        self._comm_stream = None

    @property
    def angles(self):
        return self.base.angles

    # ---------------------------------------------------------------- A ----
    def trafo(self, x: Tensor) -> Tensor:
        return self.base._fp(x, angle_range=self.angle_range)

    __call__ = trafo

    def gather_sinogram(self, y_local: Tensor) -> Tensor:
        """Full sinogram on every rank (rows are disjoint, so a sum all-reduce assembles them)."""
        if self.world > 1:
            y_local = y_local.clone()
            dist.all_reduce(y_local, op=dist.ReduceOp.SUM, group=self.group)
        return y_local

    # --------------------------------------------------------------- A* ----
    def _reduce_chunks(self, produce, n: int, out: Tensor) -> Tensor:
        """``out[c] = all_reduce(produce(c))`` over slice chunks, communication of chunk c
        overlapping the computation of chunk c+1 (side stream on CUDA)."""
        if self.world == 1:
            for lo in range(0, n, self.chunk):
                out[lo:lo + self.chunk] = produce(lo, min(n, lo + self.chunk))
            return out
        on_cuda = out.is_cuda
        if on_cuda and self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=out.device)
        pending = []
        for lo in range(0, n, self.chunk):
            hi = min(n, lo + self.chunk)
            part = produce(lo, hi)
            if on_cuda:
                ready = torch.cuda.Event()
                ready.record(torch.cuda.current_stream(out.device))
                with torch.cuda.stream(self._comm_stream):
                    self._comm_stream.wait_event(ready)
                    dist.all_reduce(part, op=dist.ReduceOp.SUM, group=self.group)
                    part.record_stream(self._comm_stream)
                pending.append((lo, hi, part))
            else:
                work = dist.all_reduce(part, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
                pending.append((lo, hi, part, work))
        if on_cuda:
            torch.cuda.current_stream(out.device).wait_stream(self._comm_stream)
            for lo, hi, part in pending:
                out[lo:hi] = part
        else:
            for lo, hi, part, work in pending:
                work.wait()
                out[lo:hi] = part
        return out

    def trafo_adjoint(self, y: Tensor) -> Tensor:
        lead = y.shape[:-2]
        yf = y.reshape(-1, 1, *self.obs_shape)
        out = torch.empty(yf.shape[0], 1, *self.im_shape, dtype=y.dtype, device=y.device)
        self._reduce_chunks(
            lambda lo, hi: self.base._bp(yf[lo:hi], self.base.adj_scale, angle_range=self.angle_range),
            yf.shape[0], out)
        return out.reshape(*lead, *self.im_shape)

    def normal_apply(self, v: Tensor, gamma: float) -> Tensor:
        lead = v.shape[:-2]
        vf = v.reshape(-1, 1, *self.im_shape)
        out = torch.empty_like(vf)

        fused = getattr(self.base, 'normal_apply', None)      # B200RayTrafo: A*A without re-laying-out the sinogram

        def produce(lo, hi):
            if fused is not None:
                return fused(vf[lo:hi], gamma, angle_range=self.angle_range, add_identity=False)
            q = self.base._fp(vf[lo:hi], angle_range=self.angle_range)
            return self.base._bp(q, gamma * self.base.adj_scale, angle_range=self.angle_range)
        self._reduce_chunks(produce, vf.shape[0], out)
        return (vf + out).reshape(*lead, *self.im_shape)

    def normal_op(self, gamma: float):
        return lambda v: self.normal_apply(v, gamma)
