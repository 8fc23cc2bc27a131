"""Thin wrappers over the geometry-free C-ABI entry points (Tweedie, DDIM)."""
import ctypes as C

import torch
from torch import Tensor

from . import _lib


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _check_cuda_f32(**tensors):
    dev = None
    for name, t in tensors.items():
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError('%s must be a CUDA tensor (no CPU fallback in libscd_b200)' % name)
        if t.dtype != torch.float32:
            raise TypeError('%s must be float32, got %s' % (name, t.dtype))
        dev = dev or t.device
        if t.device != dev:
            raise ValueError('%s is on %s, expected %s' % (name, t.device, dev))
    return dev


def _times(t: Tensor, batch: int, device) -> Tensor:
    t = t.to(device=device, dtype=torch.float32).reshape(-1).contiguous()
    if t.numel() != batch:
        raise ValueError('time step tensor must have one entry per sample (%d != %d)' % (t.numel(), batch))
    return t


def tweedie_rhs(x: Tensor, s: Tensor, t: Tensor, abar: Tensor, atb: Tensor = None, gamma: float = 0.0):
    """``xhat0 = (x - s*std_t)/mean_t`` and, if ``atb`` is given, ``b = xhat0 + gamma*atb``."""
    dev = _check_cuda_f32(x=x, s=s, atb=atb, abar=abar)
    x = x.contiguous(); s = s.contiguous()
    if s.shape != x.shape:
        raise ValueError('x %r and s %r differ in shape' % (tuple(x.shape), tuple(s.shape)))
    batch = x.shape[0]
    numel = x[0].numel()
    t = _times(t, batch, dev)
    xhat0 = torch.empty_like(x)
    b = None
    atb_ptr = None
    if atb is not None:
        atb = atb.expand_as(x).contiguous()
        b = torch.empty_like(x)
        atb_ptr = atb.data_ptr()
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(lib.scd_tweedie_rhs(x.data_ptr(), s.data_ptr(), atb_ptr, t.data_ptr(), abar.data_ptr(),
                                       int(abar.numel()), float(gamma), xhat0.data_ptr(),
                                       b.data_ptr() if b is not None else None, batch, numel, _stream(dev)),
                   'scd_tweedie_rhs')
    return (xhat0, b) if atb is not None else xhat0


def ddim_ddpm(xhat: Tensor, s: Tensor, eps: Tensor, t: Tensor, t_prev: Tensor, abar: Tensor, eta: float):
    """DDPM branch of the reference's ``ddim`` as one kernel."""
    dev = _check_cuda_f32(xhat=xhat, s=s, eps=eps, abar=abar)
    xhat = xhat.contiguous(); s = s.contiguous(); eps = eps.contiguous()
    if s.shape != xhat.shape or eps.shape != xhat.shape:
        raise ValueError('xhat, s and eps must have the same shape')
    batch = xhat.shape[0]
    numel = xhat[0].numel()
    t = _times(t, batch, dev)
    t_prev = _times(t_prev, batch, dev)
    out = torch.empty_like(xhat)
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(lib.scd_ddim(xhat.data_ptr(), s.data_ptr(), eps.data_ptr(), t.data_ptr(), t_prev.data_ptr(),
                                abar.data_ptr(), int(abar.numel()), float(eta), out.data_ptr(), batch, numel,
                                _stream(dev)), 'scd_ddim')
    return out


def launch_count(reset=False):
    lib = _lib.load()
    n = int(lib.scd_launch_count())
    if reset:
        lib.scd_launch_count_reset()
    return n
