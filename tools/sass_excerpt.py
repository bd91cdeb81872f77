"""SASS evidence for DESIGN.md (no GPU needed): opcode histograms of the main kernels and excerpts of the two inner
loops the design argues about, from `cuobjdump -sass` of the built library.

    python tools/sass_excerpt.py > profiles/sass_r02.md
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'swiftortho_b200', 'libswiftortho_b200.so')
lines = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout.split('\n')


def func_range(name):
    st = [i for i, l in enumerate(lines) if 'Function :' in l and name in l][0]
    en = next((i for i in range(st + 1, len(lines)) if 'Function :' in lines[i]), len(lines))
    return st, en


def instrs(st, en):
    out = []
    for l in lines[st:en]:
        m = re.search(r'/\*([0-9a-f]{4})\*/\s+((?:@!?U?P\d\s+)?[A-Z0-9_.]+.*?);', l)
        if m:
            out.append((m.group(1), m.group(2).strip()))
    return out


def opcode(text):
    return re.sub(r'^@!?U?P\d\s+', '', text).split()[0].split('.')[0]


print('# SASS excerpts, round 2 (final build)\n')
print('`cuobjdump -sass swiftortho_b200/libswiftortho_b200.so` (sm_100a cubins, CUDA 12.9), summarised by '
      '`tools/sass_excerpt.py`.  DESIGN.md 4.1 / 4.3 argue from these instruction mixes: an X-drop step is '
      '`VIADDMNMX` + `ISETP` + `PRMT` on the alu pipe, `IMAD` forms on the fma pipe and one `LDS`; a banded-DP cell uses '
      'the DPX min/max forms (`VIMNMX3`, `VIADDMNMX.RELU`, `VIMNMX.RELU`) with `PRMT`-spliced table addresses.  No kernel '
      'contains a tensor-core or TMA instruction (`UTCMMA`, `UTMALDG`, `LDTM`): no stage of this path is a dense '
      'contraction (BASELINE.json north_star).\n')
for name, title in (('k_xdropILi24ELb1ELi32ELi16', 'k_xdrop<24, FAST, 32 warps, 16 query views> (default layout)'),
                    ('k_xdropILi24ELb1ELi16ELi0', 'k_xdrop<24, FAST, 16 warps, 1 query view> (first half of round 2)'),
                    ('k_cell_span', 'k_cell_span'), ('k_banded_dp_wave', 'k_banded_dp_wave (16 lanes per alignment)'),
                    ('k_banded_dpEPK', 'k_banded_dp'), ('k_mcl_spgemm_warpILb1', 'k_mcl_spgemm_warp<write>'),
                    ('k_apc_rows', 'k_apc_rows'),
                    ('k_traceback', 'k_traceback'), ('k_h3_select', 'k_h3_select (H3 on the device)'), ('k_cand_sort', 'k_cand_sort'),
                    ('k_cell_passILb1', 'k_cell_pass<true>'), ('k_cell_small', 'k_cell_small'), ('k_orth_classify', 'k_orth_classify')):
    st, en = func_range(name)
    ins = instrs(st, en)
    c = collections.Counter(opcode(t) for _, t in ins)
    full = collections.Counter(re.sub(r'^@!?U?P\d\s+', '', t).split()[0] for _, t in ins)
    print('## %s: %d SASS instructions\n' % (title, len(ins)))
    print('opcodes: ' + ', '.join('%s %d' % kv for kv in c.most_common(18)) + '\n')
    dpx = sorted(k for k in full if k.startswith(('VIADDMNMX', 'VIMNMX')))
    tens = sum(v for k, v in full.items() if k.startswith(('UTMA', 'UTC', 'LDTM', 'HMMA', 'IMMA')))
    print('min/max (DPX) forms: ' + (', '.join('%s %d' % (k, full[k]) for k in dpx) or 'none') + '; tensor / TMA forms: %d\n' % tens)


def excerpt(name, pred, width, title):
    st, en = func_range(name)
    ins = instrs(st, en)
    hits = [i for i, (_, t) in enumerate(ins) if pred(t)]
    best, lo = 0, 0
    for i in hits:                                   # densest window
        n = sum(1 for j in hits if i <= j < i + width)
        if n > best:
            best, lo = n, i
    print('## %s\n\n```' % title)
    for a, t in ins[max(0, lo - 2):lo + width]:
        print('/*%s*/  %s ;' % (a, t))
    print('```\n')


excerpt('k_xdropILi24ELb1ELi32ELi16', lambda t: 'VIADDMNMX' in t, 26,
        'k_xdrop<24, FAST>: consecutive extension steps of the unrolled 16-step body (VIADDMNMX = d = max(d + e, 0), '
        'predicated IMAD forms = the v / d updates on the fma pipe, ISETP = X-drop test)')
excerpt('k_banded_dp', lambda t: 'VIMNMX3' in t or 'VIADDMNMX' in t, 30,
        'k_banded_dp: cells of the band loop (VIADDMNMX.RELU / VIMNMX3 = max(0, I, M, D), PRMT = table address and trace code splice)')
