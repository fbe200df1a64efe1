"""Instruction / stall-sample shares of the phases of descend_fx_kernel from an ncu source page + nvdisasm -g listing:
python tools/sass_regions_fx.py page.csv disasm.sass kernel_substring"""
import collections, csv, re, sys
page, sass, kname = sys.argv[1:4]
src = open('boardlaw_b200/csrc/descend_fx.cu').read().split('\n')
marks = [(i + 1, l.strip()[:60]) for i, l in enumerate(src) if re.match(r'\s*// --', l)]
kstart = next(i + 1 for i, l in enumerate(src) if 'descend_fx_kernel(' in l)
kend = next(i + 1 for i, l in enumerate(src) if l.startswith('int g_fx_nit'))
bounds = [(0, kstart, 'helpers (inlined, unattributed)'), (kstart, marks[0][0], 'prologue')]
for (a, n), (b, _) in zip(marks, marks[1:] + [(kend, '')]):
    bounds.append((a, b, n))
rows = list(csv.reader(open(page)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
h = rows[hi]; ci, ct, cst = h.index('Instructions Executed'), h.index('Thread Instructions Executed'), h.index('# Samples')
inst = [(int(r[ci]), int(r[ct]), int(r[cst])) for r in rows[hi + 1:] if len(r) > ct]
lines, cur, on = [], None, False
for l in open(sass):
    if l.startswith('//---') and '.text.' in l:
        on = kname in l; continue
    if not on: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)), m.group(3)); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: lines.append((cur, m.group(2)))
def region(loc):
    if loc is None: return 'unknown'
    f, ln, extra = loc
    # innermost frame that lies in the kernel body
    cands = [ln] if f == 'descend_fx.cu' else []
    cands += [int(x) for x in re.findall(r'inlined at "[^"]*descend_fx.cu", line (\d+)', extra or '')]
    for c in cands:
        if c >= kstart:
            for a, b, n in bounds:
                if a <= c < b: return n
    return 'other: ' + f
reg, regs, regt = collections.Counter(), collections.Counter(), collections.Counter()
for (n, tn, s), (loc, txt) in zip(inst, lines):
    r = region(loc); reg[r] += n; regs[r] += s; regt[r] += tn
tot, tots = sum(reg.values()), sum(regs.values())
print(f'{tot} warp instructions, {tots} samples')
for k, v in reg.most_common():
    print(f'{100 * v / tot:5.1f}% inst {100 * regs[k] / max(tots, 1):5.1f}% samples  lanes/inst {regt[k] / max(v, 1):5.1f}  {k}')
