"""Per-kernel SASS instruction histogram of libboardlaw_b200.so (no GPU needed): counts of the mnemonics that prove which hardware
path a kernel uses — tcgen05 MMAs (UTCHMMA), tensor-memory loads/stores (LDTM/STTM), TMA-engine bulk copies (UBLKCP), cp.async
(LDGSTS), packed fp32 (FFMA2/FMUL2/FADD2), 128-bit global loads — and the kernel's size.
python tools/sass_hist.py [lib.so] > profiles/r02_sass_hist.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else 'boardlaw_b200/libboardlaw_b200.so'
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
KEYS = ['UTCHMMA', 'LDTM', 'STTM', 'UBLKCP', 'UTMALDG', 'LDGSTS', 'FFMA2', 'FMUL2', 'FADD2', 'LDG.E.128', 'STG.E.128', 'MUFU.RCP', 'DADD', 'DFMA', 'SHFL', 'BAR.SYNC', 'SYNCS']
kern, hist, total = None, collections.defaultdict(collections.Counter), collections.Counter()
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        kern = m.group(1)
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and kern:
        op = m.group(1)
        total[kern] += 1
        for k in KEYS:
            if op.startswith(k):
                hist[kern][k] += 1
names = subprocess.run(['c++filt'], input='\n'.join(total), capture_output=True, text=True).stdout.splitlines()
short = lambda n: re.sub(r'\(.*', '', re.sub(r'\(anonymous namespace\)::|<unnamed>::|^void ', '', n))
print(f'{"kernel":58s} {"SASS":>6s}  ' + ' '.join(f'{k:>9s}' for k in KEYS))
rows = sorted(zip(names, total), key=lambda x: short(x[0]))
for name, k in rows:
    print(f'{short(name)[:58]:58s} {total[k]:6d}  ' + ' '.join(f'{hist[k][x] or "":>9}' for x in KEYS))
print('\ntotals: ' + ', '.join(f'{x} {sum(h[x] for h in hist.values())}' for x in KEYS if sum(h[x] for h in hist.values())))
