"""Out-of-bounds write check of every C-ABI entry point (compute-sanitizer is not available on the GPU pool).

Every device buffer handed to the library -- outputs, scratch, workspaces, exactly as large as the
``*_bytes`` queries say -- sits between two guard regions filled with a byte pattern; after the calls the
guards must be untouched and every output element must have been written (the outputs start as NaN)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GUARD = 8192
PATTERN = 0xA5


class Guarded:
    """``nbytes`` usable bytes (256-byte aligned) with GUARD pattern bytes on either side."""

    def __init__(self, nbytes, dev, fill_nan=True):
        self.nbytes = int(nbytes)
        self.raw = torch.full((GUARD + self.nbytes + GUARD,), PATTERN, dtype=torch.uint8, device=dev)
        assert (self.raw.data_ptr() + GUARD) % 256 == 0
        if fill_nan and self.nbytes:
            self.raw[GUARD:GUARD + self.nbytes] = 0xFF          # 0xFFFFFFFF = NaN as fp32
        self.ptr = self.raw.data_ptr() + GUARD

    def floats(self, *shape):
        n = int(np.prod(shape))
        assert n * 4 <= self.nbytes
        return self.raw[GUARD:GUARD + n * 4].view(torch.float32).reshape(*shape)

    def set(self, t):
        self.floats(*t.shape).copy_(t)
        return self

    def intact(self):
        lo, hi = self.raw[:GUARD], self.raw[GUARD + self.nbytes:]
        return bool((lo == PATTERN).all()) and bool((hi == PATTERN).all())


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


@pytest.mark.parametrize('im_shape,num_angles,batches', [
    ((256, 256), 60, (1, 3, 8, 9, 17, 33)),
    ((33, 47), 5, (1, 2, 5)),
    ((501, 501), 24, (2,)),
    ((64, 200), 7, (4,)),
])
def test_no_write_outside_the_buffers(im_shape, num_angles, batches):
    import diffusion_models_dev_project_b200 as pkg
    from diffusion_models_dev_project_b200 import _lib
    dev = torch.device('cuda')
    lib = _lib.load()
    rt = pkg.B200RayTrafo(im_shape, num_angles)
    h = rt._handle(dev)
    na, nd = rt.obs_shape
    n_img, n_sino = im_shape[0] * im_shape[1], na * nd
    abar = pkg.DDPM().alpha_bar_table(dev)
    st = _stream(dev)
    for B in batches:
        torch.manual_seed(B)
        x = torch.rand(B, 1, *im_shape, device=dev)
        ref_y = rt(x)
        ref_z = rt.trafo_adjoint(ref_y)
        bufs = {}

        def G(name, nbytes, **kw):
            bufs[name] = Guarded(nbytes, dev, **kw)
            return bufs[name]

        img = G('img', B * n_img * 4).set(x)
        # ---- A, full and partial angle range
        sino = G('sino', B * n_sino * 4)
        scr = G('fp_scratch', lib.scd_fp_scratch_bytes(h.ptr, B))
        _lib.check(lib.scd_fp(h.ptr, img.ptr, sino.ptr, B, 0, na, scr.ptr, scr.nbytes, st), 'scd_fp')
        y = sino.floats(B, 1, na, nd)
        assert torch.equal(y, ref_y)
        if na >= 3:
            sino2 = G('sino_part', B * n_sino * 4)
            sino2.floats(B, 1, na, nd).zero_()
            _lib.check(lib.scd_fp(h.ptr, img.ptr, sino2.ptr, B, 1, na - 1, scr.ptr, scr.nbytes, st), 'scd_fp range')
            y2 = sino2.floats(B, 1, na, nd)
            assert torch.equal(y2[:, :, 1:na - 1], ref_y[:, :, 1:na - 1]) and float(y2[:, :, 0].abs().max()) == 0.0
        # ---- A*, plain and with the fused addend
        out = G('bp_out', B * n_img * 4)
        bscr = G('bp_scratch', lib.scd_bp_scratch_bytes(h.ptr, B))
        _lib.check(lib.scd_bp(h.ptr, sino.ptr, out.ptr, B, 0, na, rt.adj_scale, None, 0.0, bscr.ptr, bscr.nbytes, st), 'scd_bp')
        assert torch.equal(out.floats(B, 1, *im_shape), ref_z)
        out2 = G('bp_out_add', B * n_img * 4)
        _lib.check(lib.scd_bp(h.ptr, sino.ptr, out2.ptr, B, 0, na, 0.5, img.ptr, 2.0, bscr.ptr, bscr.nbytes, st), 'scd_bp addend')
        assert torch.isfinite(out2.floats(B, 1, *im_shape)).all()
        # ---- the pair on the interleaved sinogram
        il = G('sino_il', lib.scd_sino_il_buffer_bytes(h.ptr, B))
        _lib.check(lib.scd_fp_il(h.ptr, img.ptr, il.ptr, B, 0, na, scr.ptr, scr.nbytes, st), 'scd_fp_il')
        out3 = G('bp_il_out', B * n_img * 4)
        _lib.check(lib.scd_bp_il(h.ptr, il.ptr, out3.ptr, B, 0, na, rt.adj_scale, None, 0.0, st), 'scd_bp_il')
        assert torch.equal(out3.floats(B, 1, *im_shape), ref_z)
        # ---- CG and the fused DDS step
        work = G('cg_work', lib.scd_cg_workspace_bytes(h.ptr, B))
        xs = G('cg_x', B * n_img * 4).set(x)
        rhs = G('cg_rhs', B * n_img * 4).set(ref_z)
        _lib.check(lib.scd_cg(h.ptr, xs.ptr, rhs.ptr, 0.05, 2, B, work.ptr, work.nbytes, st), 'scd_cg')
        assert torch.isfinite(xs.floats(B, 1, *im_shape)).all()
        s = G('score', B * n_img * 4).set(torch.randn(B, 1, *im_shape, device=dev))
        eps = G('eps', B * n_img * 4).set(torch.randn(B, 1, *im_shape, device=dev))
        t = G('t', B * 4).set(torch.full((B,), 500., device=dev))
        tp = G('tp', B * 4).set(torch.full((B,), 490., device=dev))
        xn, xh = G('x_next', B * n_img * 4), G('xhat0', B * n_img * 4)
        _lib.check(lib.scd_dds_step(h.ptr, img.ptr, s.ptr, rhs.ptr, eps.ptr, t.ptr, tp.ptr, abar.data_ptr(), abar.numel(),
                                    0.05, 0.15, 2, xn.ptr, xh.ptr, B, work.ptr, work.nbytes, st), 'scd_dds_step')
        assert torch.isfinite(xn.floats(B, n_img)).all() and torch.isfinite(xh.floats(B, n_img)).all()
        # ---- vector kernels
        tw, tb = G('tweedie_out', B * n_img * 4), G('tweedie_b', B * n_img * 4)
        _lib.check(lib.scd_tweedie_rhs(img.ptr, s.ptr, rhs.ptr, t.ptr, abar.data_ptr(), abar.numel(), 0.05, tw.ptr, tb.ptr,
                                       B, n_img, st), 'scd_tweedie_rhs')
        dd = G('ddim_out', B * n_img * 4)
        _lib.check(lib.scd_ddim(tw.ptr, s.ptr, eps.ptr, t.ptr, tp.ptr, abar.data_ptr(), abar.numel(), 0.15, dd.ptr, B, n_img, st),
                   'scd_ddim')
        assert torch.isfinite(tw.floats(B, n_img)).all() and torch.isfinite(tb.floats(B, n_img)).all()
        assert torch.isfinite(dd.floats(B, n_img)).all()
        # ---- adaptation-loss kernels
        res = G('residual', B * n_sino * 4)
        rpart = G('residual_part', lib.scd_residual_sq_blocks(B * n_sino) * 4)
        _lib.check(lib.scd_residual_sq(sino.ptr, sino.ptr, res.ptr, rpart.ptr, B * n_sino, st), 'scd_residual_sq')
        assert float(res.floats(B * n_sino).abs().max()) == 0.0 and float(rpart.floats(rpart.nbytes // 4).abs().max()) == 0.0
        tvp = G('tv_part', B * lib.scd_tv_blocks(*im_shape) * 4)
        tvg = G('tv_grad', B * n_img * 4)
        _lib.check(lib.scd_tv_loss(img.ptr, tvp.ptr, B, im_shape[0], im_shape[1], st), 'scd_tv_loss')
        _lib.check(lib.scd_tv_grad(img.ptr, tvg.ptr, B, im_shape[0], im_shape[1], st), 'scd_tv_grad')
        assert torch.isfinite(tvp.floats(tvp.nbytes // 4)).all() and torch.isfinite(tvg.floats(B, n_img)).all()
        assert abs(float(tvp.floats(tvp.nbytes // 4).sum()) - float(pkg.tv_loss(x))) <= 1e-4 * float(pkg.tv_loss(x))
        # ---- ramp filter of fbp
        filt = G('ramp_out', B * n_sino * 4)
        _lib.check(lib.scd_ramp_filter(h.ptr, sino.ptr, filt.ptr, B, st), 'scd_ramp_filter')
        assert torch.equal(filt.floats(B, 1, na, nd), rt.ramp_filter(ref_y))
        # ---- sample-interleaved images (batches >= 3)
        n_il = lib.scd_img_il_bytes(h.ptr, B)
        assert (n_il > 0) == (B >= 3 or (B == 1 and im_shape[1] % 4 == 0))
        if n_il:
            x_il, z_il = G('img_il', n_il), G('bp_il_img', n_il)
            _lib.check(lib.scd_img_il_pack(h.ptr, img.ptr, x_il.ptr, B, st), 'scd_img_il_pack')
            back = G('img_unpacked', B * n_img * 4)
            _lib.check(lib.scd_img_il_unpack(h.ptr, x_il.ptr, back.ptr, B, st), 'scd_img_il_unpack')
            assert torch.equal(back.floats(B, 1, *im_shape), x)
            il2 = G('sino_il2', lib.scd_sino_il_buffer_bytes(h.ptr, B))
            _lib.check(lib.scd_fp_ilimg(h.ptr, x_il.ptr, il2.ptr, B, 0, na, st), 'scd_fp_ilimg')
            _lib.check(lib.scd_bp_ilimg(h.ptr, il2.ptr, z_il.ptr, B, 0, na, rt.adj_scale, x_il.ptr, 0.0, st), 'scd_bp_ilimg')
            _lib.check(lib.scd_img_il_unpack(h.ptr, z_il.ptr, back.ptr, B, st), 'scd_img_il_unpack')
            assert torch.equal(back.floats(B, 1, *im_shape), ref_z)
        else:
            assert lib.scd_img_il_pack(h.ptr, img.ptr, img.ptr, B, st) == _lib.SCD_E_INVALID
        # ---- SCD adaptation objective: forward, reverse sweep (all three data-consistency types)
        loss, gs = G('adapt_loss', 256), G('adapt_grad', B * n_img * 4)
        for dc, k in ((0, 2), (1, 1), (2, 1)):
            aw = G('adapt_work_%d' % dc, lib.scd_adapt_workspace_bytes(h.ptr, B, k))
            _lib.check(lib.scd_adapt_fwd(h.ptr, img.ptr, s.ptr, rhs.ptr, sino.ptr, t.ptr, abar.data_ptr(), abar.numel(), 0.05, k, dc,
                                         1e-3, loss.ptr, None, B, aw.ptr, aw.nbytes, st), 'scd_adapt_fwd')
            _lib.check(lib.scd_adapt_bwd(h.ptr, None, t.ptr, abar.data_ptr(), abar.numel(), 0.05, k, dc, 1e-3,
                                         rt.adj_scale / rt.geometry.range_weight, gs.ptr, B, aw.ptr, aw.nbytes, st), 'scd_adapt_bwd')
            assert torch.isfinite(gs.floats(B, n_img)).all() and bool(torch.isfinite(loss.floats(1)).all())
        torch.cuda.synchronize()
        broken = [k for k, g in bufs.items() if not g.intact()]
        assert not broken, 'guard bytes overwritten around %s (shape %r, batch %d)' % (broken, im_shape, B)
        # inputs are read-only
        assert torch.equal(img.floats(B, 1, *im_shape), x)
