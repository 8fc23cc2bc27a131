"""Per-CTA phase timeline of fp_march / bp_tile (%globaltimer stamps, scd_debug_set_stamps).

Needs the debug build of the library (compiled with -DSCD_DEBUG_STAMPS; built here if missing -- needs nvcc):
the default libscd_b200.so carries no stamp code.

    python tools/timeline.py --kernel fp --batch 8
Prints, per phase boundary, min / median / max over the CTAs of the time since the earliest CTA start.
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ['SCD_B200_LIB'] = os.path.join(_ROOT, 'diffusion_models_dev_project_b200', '_lib', 'libscd_b200_dbg.so')   # before the import
import diffusion_models_dev_project_b200 as pkg  # noqa: E402
from diffusion_models_dev_project_b200 import build as _build  # noqa: E402
assert _build.build(debug=True) == os.environ['SCD_B200_LIB']
from diffusion_models_dev_project_b200 import _lib  # noqa: E402

NAMES = {'fp': ['cta start', 'tables done', 'predecessor complete', 'first strip landed', 'march done',
                'partials exchanged', 'output written', 'exit'],
         'bp': ['cta start', 'tables done', 'predecessor complete', 'first chunk landed', 'warp0 march done',
                'epilogue done', '-', '-']}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--kernel', default='fp', choices=['fp', 'bp'])
    ap.add_argument('--batch', type=int, default=8)
    ap.add_argument('--im', type=int, default=256)
    ap.add_argument('--angles', type=int, default=60)
    ap.add_argument('--set', default='')
    ap.add_argument('--dump', default='', help='write the raw stamps (ns, [CTA, 8]) to this .npy file')
    ap.add_argument('--mod', type=int, default=0, help='also print the median CTA duration per (CTA index mod MOD): with MOD = units per group of fp_march it separates the angle classes')
    a = ap.parse_args()
    dev = torch.device('cuda')
    rt = pkg.B200RayTrafo((a.im, a.im), a.angles)
    tune = {k: int(v) for k, v in (kv.split('=') for kv in a.set.split(',') if kv)}
    if tune:
        rt.set_tuning(dev, **tune)
    lib = _lib.load()
    import ctypes
    lib.scd_debug_set_stamps.argtypes = [ctypes.c_void_p]
    lib.scd_debug_set_stamps.restype = None
    x = torch.rand(a.batch, 1, a.im, a.im, device=dev)
    p = torch.rand_like(x)
    q = rt._fp_il(x)
    lead = x.shape[:-2]
    fn = (lambda: rt._fp_il(x)) if a.kernel == 'fp' else (lambda: rt._bp_il(q, lead, 0.01, addend=p, addend_scale=1.0))
    for _ in range(3):
        fn()
    stamps = torch.zeros(1 << 20, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    lib.scd_debug_set_stamps(stamps.data_ptr())
    fn()
    torch.cuda.synchronize()
    lib.scd_debug_set_stamps(None)
    s = stamps.cpu().numpy().reshape(-1, 8)
    if a.dump:
        np.save(a.dump, s[:4096])
    s = s[s[:, 0] > 0]
    t0 = s[:, 0].min()
    print('%s at batch %d: %d CTAs, span %.1f us' % (a.kernel, a.batch, len(s), (s[:, :7].max() - t0) / 1e3))
    for k, name in enumerate(NAMES[a.kernel][:7] if a.kernel == 'fp' else NAMES[a.kernel]):
        col = s[:, k]
        col = col[col > 0]
        if len(col) == 0:
            continue
        rel = (col - t0) / 1e3
        print('  %-22s min %7.2f  med %7.2f  max %7.2f us' % (name, rel.min(), np.median(rel), rel.max()))
    dur = (s[:, [c for c in range(7 if a.kernel == 'fp' else 8) if (s[:, c] > 0).all()][-1]] - s[:, 0]) / 1e3
    print('  per-CTA duration       min %7.2f  med %7.2f  max %7.2f us' % (dur.min(), np.median(dur), dur.max()))
    if a.kernel == 'fp':
        tag = s[:, 7]
        march = (s[:, 4] - s[:, 3]) / 1e3
        total = (s[:, 6] - s[:, 0]) / 1e3
        print('  by (class, angles per unit): CTAs, median duration, median march phase')
        for t in sorted(set(tag.tolist())):
            m = tag == t
            print('    class %d, %d angle(s): %4d CTAs  %7.1f us  march %7.1f us' % (t // 1000, t % 1000, m.sum(),
                                                                                np.median(total[m]), np.median(march[m])))
    if a.mod:
        idx = np.arange(len(dur))
        march = (s[:, 4] - s[:, 3]) / 1e3
        print('  median duration / march phase per CTA index mod %d:' % a.mod)
        print('   ', ' '.join('%d:%.0f/%.0f' % (m, np.median(dur[idx % a.mod == m]), np.median(march[idx % a.mod == m])) for m in range(a.mod)))


if __name__ == '__main__':
    main()
