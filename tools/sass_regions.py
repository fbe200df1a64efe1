"""Instruction and stall-sample shares of the phases of descend_mw_kernel (ncu SASS page + nvdisasm line info).
python tools/sass_regions.py page.csv disasm.sass kernel_substring n_warps"""
import csv, re, sys
page, sass, kname, W = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
rows = list(csv.reader(open(page)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address'); h = rows[hi]
ci, cs, ct, cm = h.index('Instructions Executed'), h.index('Source'), h.index('Thread Instructions Executed'), h.index('# Samples')
inst = [(r[cs].strip(), int(r[ci]), int(r[ct]), int(r[cm])) for r in rows[hi + 1:] if len(r) > ct]
lines, cur, on = [], None, False
for l in open(sass):
    if l.startswith('//---') and '.text.' in l:
        on = kname in l; continue
    if not on: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: lines.append((cur, m.group(2)))
src = open('boardlaw_b200/csrc/descend_mw.cu').read().split('\n')
def find(s): return next(i + 1 for i, l in enumerate(src) if l.strip().startswith(s))
marks = [(find('while (true) {'), 'loop head + gate'), (find('if (state == ST_SAMPLE)'), 'sample'), (find('if (state == ST_ADVANCE)'), 'advance'),
         (find('if (state == ST_DONE)'), 'done + wait'), (find('if (visit) {'), 'visit'), (find('bool pass = state'), 'terms'),
         (find('bool bad = false;'), 'child terms'), (find('const bool slow = state'), 'slow path'), (find('pass = state == ST_PASS'), 'chain'),
         (find('const float accS'), 'newton'), (find('// ---- expand + env step of the group'), 'tail')]
first = marks[0][0]
agg, last = {}, 0
for (op, n, tn, sm), (loc, txt) in zip(inst, lines):
    if loc and loc[0] == 'descend_mw.cu' and loc[1] >= first: last = loc[1]
    name = 'prologue'
    for ln, nm in marks:
        if last >= ln: name = nm
    a = agg.setdefault(name, [0, 0, 0, 0]); a[0] += n; a[1] += tn; a[2] += 1; a[3] += sm
tot, tots = sum(a[0] for a in agg.values()), sum(a[3] for a in agg.values())
print(f'{tot} warp instructions, {tot / W:.0f} per warp; {tots} samples')
for name in ['prologue'] + [m[1] for m in marks]:
    a = agg.get(name, [0, 0, 0, 0])
    print(f'{name:16s} static {a[2]:5d}  per warp {a[0] / W:8.0f} ({100 * a[0] / tot:4.1f}%)  thr/inst {a[1] / max(a[0], 1):5.1f}  samples {100 * a[3] / max(tots, 1):5.1f}%  cycles/inst {a[3] / max(a[0], 1) * tot / max(tots, 1):5.2f} (rel)')
