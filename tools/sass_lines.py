"""Per-source-line instruction counts of one kernel: joins an ncu SASS source page (csv) with nvdisasm -g line info.
python tools/sass_lines.py page.csv disasm.sass kernel_substring source.cu"""
import csv, re, sys
page, sass, kname, srcfile = sys.argv[1:5]
rows = list(csv.reader(open(page)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
h = rows[hi]
ci, cs, ct, cst = h.index('Instructions Executed'), h.index('Source'), h.index('Thread Instructions Executed'), h.index('# Samples')
inst = [(r[cs].strip(), int(r[ci]), int(r[ct]), int(r[cst])) for r in rows[hi + 1:] if len(r) > ct]
# disassembly: sequence of instructions with current (file,line)
lines, cur, on = [], None, False
for l in open(sass):
    if l.startswith('//---') and '.text.' in l:
        on = kname in l
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        lines.append((cur, m.group(2)))
print(len(inst), 'profiled instructions,', len(lines), 'disassembled')
agg = {}
for (op, n, tn, smp), (loc, txt) in zip(inst, lines):
    a = agg.setdefault(loc, [0, 0, 0]); a[0] += n; a[1] += tn; a[2] += smp
tot = sum(a[0] for a in agg.values()); tots = sum(a[2] for a in agg.values())
src = open(srcfile).read().split('\n')
base = srcfile.split('/')[-1]
print(f'total warp instructions {tot}, samples {tots}')
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[5]) if len(sys.argv) > 5 else 45]:
    text = src[loc[1] - 1].strip()[:90] if loc and loc[0] == base and loc[1] <= len(src) else ''
    print(f'{100*a[0]/tot:5.1f}% inst {100*a[2]/max(tots,1):5.1f}% smp  thr/inst {a[1]/max(a[0],1):5.1f}  {loc}  {text}')
