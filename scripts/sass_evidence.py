#!/usr/bin/env python
"""Regenerate profiles/sass_*.txt: the Blackwell-native instructions in the built libdreamgnn.so, per kernel.

    python scripts/sass_evidence.py            (after `python -m dreamgnn_b200.build`; no GPU needed)

For each kernel family the file lists, per SASS function, the total instruction count, the counts of the mnemonics that
prove the tcgen05 / TMEM / TMA / L2-prefetch paths (B200_PROFILING.md "What proves a Blackwell-native kernel"), and the
first occurrences of each with their addresses -- excerpts of `cuobjdump -sass`, not the full listing.
"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
LIB = os.path.join(REPO, 'dreamgnn_b200', 'lib', 'libdreamgnn.so')
OUT = os.path.join(REPO, 'profiles')
MNEMONICS = ('UTCHMMA', 'UTCQMMA', 'UTCBAR', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'UBLKPF', 'LDTM', 'STTM', 'CCTL.E.PF2', 'SYNCS',
             'LDGSTS', 'STG.E.256', 'LDG.E.128', 'HMMA', 'FFMA2', 'FADD2', 'FMUL2', 'REDG', 'ATOMG')
FAMILIES = {
    'sass_gemm_tc.txt': ('gemm_nt_tf32', 'tcgen05 projection GEMM (csrc/gemm_tc.cu)'),
    'sass_decoder_tc.txt': ('decoder_', 'fused pair-gather + MLP decoder (csrc/decoder_tc.cu, csrc/decoder.cu)'),
    'sass_spmm.txt': ('spmm_', 'CSR SpMM (csrc/spmm.cu)'),
    'sass_rowops.txt': ('colsum|center_normalize|attention|leaky|bce', 'row-streaming kernels (csrc/rowops.cu, csrc/fused.cu)'),
    'sass_tail.txt': ('small_gemm|adam_|bce_|gram_loss|basis_', 'small fp32 GEMM and the loss / optimiser tail (csrc/small_gemm.cu, loss.cu, optim.cu, basis.cu)'),
}


def demangle(names):
    out = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    funcs, cur = collections.OrderedDict(), None
    for ln in sass.splitlines():
        m = re.match(r'\s*Function : (\S+)', ln)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur is not None and re.match(r'\s*/\*[0-9a-f]{4,}\*/', ln):
            funcs[cur].append(ln.rstrip())
    names = demangle(list(funcs))
    digest = subprocess.run(['sha256sum', LIB], capture_output=True, text=True).stdout.split()[0][:16]
    for fname, (pat, title) in FAMILIES.items():
        lines = ['# %s -- SASS evidence' % title,
                 '# regenerate: python scripts/sass_evidence.py   (cuobjdump -sass dreamgnn_b200/lib/libdreamgnn.so, sha256 %s...)' % digest,
                 '# tcgen05.mma -> UTC*MMA; tcgen05.ld/st -> LDTM/STTM; tcgen05.commit -> UTCBAR; TMA -> UTMALDG/UTMASTG/UBLKCP;',
                 '# prefetch.global.L2 -> CCTL.E.PF2; mbarrier -> SYNCS; st.global.v8.f32 -> STG.E.256; add/fma.f32x2 -> FADD2/FFMA2', '']
        hit = False
        for mangled, body in funcs.items():
            nice = names.get(mangled, mangled)
            short = re.sub(r'\(.*', '', nice.replace('(anonymous namespace)::', ''))
            if not re.search(pat, short):
                continue
            hit = True
            counts = collections.Counter()
            first = {}
            for ln in body:
                for mn in MNEMONICS:
                    if re.search(r'\b' + re.escape(mn), ln):
                        counts[mn] += 1
                        first.setdefault(mn, []).append(ln.strip())
            lines.append('## %s' % short)
            lines.append('   %d SASS instructions; %s' % (len(body), ', '.join('%s x%d' % kv for kv in sorted(counts.items())) or 'none of the listed mnemonics'))
            for mn in ('UTCHMMA', 'UTMALDG', 'LDTM', 'STTM', 'UTCBAR', 'CCTL.E.PF2', 'STG.E.256', 'FFMA2', 'FADD2'):
                for ln in first.get(mn, [])[:2]:
                    lines.append('      ' + re.sub(r'\s+', ' ', ln)[:150])
            lines.append('')
        if hit:
            with open(os.path.join(OUT, fname), 'w') as fh:
                fh.write('\n'.join(lines))
            print('wrote', fname)


if __name__ == '__main__':
    main()
