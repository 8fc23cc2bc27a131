"""Count, per kernel instantiation, the SASS instructions that carry the Blackwell-specific parts of the design
(cuobjdump -sass of libscd_b200.so; no GPU needed).

    python tools/sass_evidence.py > profiles/rNN_sass_evidence.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'diffusion_models_dev_project_b200', '_lib', 'libscd_b200.so')
COLS = ['UTMALDG', 'UBLKCP', 'FFMA2', 'SYNCS', 'UCGABAR_ARV', 'UCGABAR_WAIT', 'LDS.128', 'ACQBULK', 'PREEXIT', 'VOTE.ALL', 'FADD.RM',
        'LD.E', 'DFMA']
WANT = [r'fp_march_kernel<4, 2, 4, 8, 24>', r'fp_march_kernel<4, 4, 6, 4, 32>', r'fp_march_kernel<4, 4, 8, 4, 24>',
        r'fp_march_kernel<4, 4, 13, 4, 16>', r'fp_march_kernel<1, 1, 6, 8, 16>', r'bp_tile_kernel<4, 2, 1, true>',
        r'bp_tile_kernel<4, 4, 4, true>', r'bp_tile_kernel<4, 4, 4, false>', r'bp_tile_kernel<1, 1, 1, false>',
        r'il_pack_kernel<16>', r'tweedie_il_kernel<8>', r'ddim_il_kernel<8>', r'cg_update_xr_il_kernel<16>',
        r'fp_packq_kernel<1, 1>', r'sino_pack_kernel', r'ramp_filter_kernel', r'adapt_update_kernel', r'band_reduce']


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    names = subprocess.run(['c++filt'], input='\n'.join(re.findall(r'Function : (\S+)', sass)), capture_output=True, text=True).stdout.split('\n')
    blocks = re.split(r'\n\s*Function : \S+', sass)[1:]
    counts = {}
    for name, body in zip(names, blocks):
        c = collections.Counter()
        for col in COLS:
            c[col] = len(re.findall(r'\b' + re.escape(col), body))
        counts[name.replace('void ', '')] = c
    print('# SASS evidence (cuobjdump -sass libscd_b200.so, sm_100a): occurrences of the instructions that carry the')
    print('# Blackwell-specific parts of the design, per kernel instantiation.')
    print('#   UTMALDG      cp.async.bulk.tensor global->shared (TMA unit, tensor map): strips gathered from the interleaved image')
    print('#   UBLKCP       cp.async.bulk global->shared (TMA unit, 1-D): strip rows / packed strips / sinogram segments')
    print('#   SYNCS        mbarrier operations (expect_tx / arrive / try_wait) of the full/empty rings')
    print('#   FFMA2        packed fp32x2 FMA (sm_100) of the interpolation taps')
    print('#   UCGABAR_*    thread-block-cluster barrier (row-split partial sums over distributed shared memory)')
    print('#   ACQBULK/PREEXIT  griddepcontrol.wait / launch_dependents (programmatic dependent launch)')
    print('#   FADD.RM      round-down add of the mantissa-floor trick (no F2I);  DFMA: fp64 accumulation of the ramp filter')
    print('%-46s' % 'kernel' + ''.join('%13s' % c for c in COLS))
    for w in WANT:
        for name, c in counts.items():
            if w in name:
                print('%-46s' % name.split('(')[0][:46] + ''.join('%13d' % c[col] for col in COLS))
                break
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    print('%-46s' % ('all %d kernels of the library' % len(counts)) + ''.join('%13d' % tot[col] for col in COLS))


if __name__ == '__main__':
    main()
