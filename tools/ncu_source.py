"""Hot source lines of one kernel in an .ncu-rep (captured with --import-source on): stall samples and executed
warp instructions per CUDA source line.

    python tools/ncu_source.py report.ncu-rep kernel_name [top]
"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', kern, '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
agg = []
for r in rows:
    if r and r[0] == 'Line No' and '# Samples' in r:
        if hdr is not None:
            break                       # first launch only
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    try:
        if r[0] == '':
            continue                    # SASS row: already counted in its source line
        ln = int(r[0])
        smp = int(r[hdr.index('# Samples')])
        ins = int(r[hdr.index('Instructions Executed')])
    except ValueError:
        continue
    agg.append((smp, ins, ln, r[1].strip()[:110]))
tot_s = sum(a[0] for a in agg) or 1
tot_i = sum(a[1] for a in agg) or 1
print('| line | samples | %% | warp-instr | %% | source |\n|---|---|---|---|---|---|')
for smp, ins, ln, src in sorted(agg, reverse=True)[:top]:
    print('| %d | %d | %.1f | %d | %.1f | `%s` |' % (ln, smp, 100 * smp / tot_s, ins, 100 * ins / tot_i, src))
print('| total | %d | | %d | | |' % (tot_s, tot_i))
