"""DRAM traffic per launch of the engine's kernel classes at the CURRENT sources, for bench.py's roofline.traffic (run on the GPU box):

    python tools/capture_traffic.py [config]   ->  profiles/r02_traffic_<config>.json

One eager move under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,...` (every launch of the move, i.e. all 63 simulations
at c2 — the r01 figure came from one late simulation); per kernel class the mean over its launches.  The file records the sha1
of boardlaw_b200/csrc, and bench.py ignores it when the sources have changed since."""
import csv
import hashlib
import json
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
CLASSES = {'descend_expand': r'descend_v3_kernel|descend_mw_kernel|descend_fx_kernel|expand_step_kernel', 'net': r'fc_tc_kernel|fc_tc_wide_kernel|fc_layer_kernel|fc_heads',
           'backup': r'backup_kernel', 'root': r'root_kernel', 'hex_step': r'hex_step_kernel|hex_transition', 'set_eval': r'set_eval_kernel', 'reset': r'reset_kernel'}


def source_sha():
    h = hashlib.sha1()
    for f in sorted((ROOT / 'boardlaw_b200' / 'csrc').glob('*.cu*')):
        h.update(f.name.encode()); h.update(f.read_bytes())
    return h.hexdigest()[:16]


def main():
    config = sys.argv[1] if len(sys.argv) > 1 else 'c2'
    out = ROOT / 'gpurun_out' / f'traffic_{config}.csv'
    out.parent.mkdir(exist_ok=True)
    metrics = 'dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,launch__registers_per_thread'
    regex = 'regex:' + '|'.join(CLASSES.values())
    subprocess.run(['ncu', '--metrics', metrics, '--clock-control', 'none', '-k', regex, '-c', '2000', '--csv', '--log-file', str(out),
                    sys.executable, str(ROOT / 'tools' / 'profile_move.py'), config, '1'], check=True, stdout=subprocess.DEVNULL, cwd=ROOT)
    rows = list(csv.reader(open(out)))
    hdr = next(r for r in rows if len(r) > 5 and r[0] == 'ID')
    per = {}
    for r in rows:
        if len(r) != len(hdr) or r[0] == 'ID':
            continue
        d = dict(zip(hdr, r))
        per.setdefault(d['ID'], {'name': d['Kernel Name']})[d['Metric Name']] = (float(d['Metric Value'].replace(',', '')), d['Metric Unit'])
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1, 'usecond': 1, 'nsecond': 1e-3, 'ms': 1e3, 'msecond': 1e3}
    agg = {}
    for d in per.values():
        cls = next((k for k, rx in CLASSES.items() if re.search(rx, d['name'])), None)
        if cls is None or 'dram__bytes_read.sum' not in d:
            continue
        val = lambda m: d[m][0] * scale.get(d[m][1], 1)
        a = agg.setdefault(cls, {'kernel': re.sub(r'\(.*', '', d['name']).replace('void ', '').replace('<unnamed>::', ''), 'launches': 0, 'dram_bytes': 0., 'us': 0.,
                                 'registers': int(d['launch__registers_per_thread'][0])})
        a['launches'] += 1
        a['dram_bytes'] += val('dram__bytes_read.sum') + val('dram__bytes_write.sum')
        a['us'] += val('gpu__time_duration.sum')
    for a in agg.values():
        a['dram_bytes_per_launch'] = a.pop('dram_bytes') / a['launches']
        a['us_per_launch_under_ncu'] = a.pop('us') / a['launches']
    dst = ROOT / 'profiles' / f'r02_traffic_{config}.json'
    dst.write_text(json.dumps({'config': config, 'source_sha': source_sha(), 'how': 'ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over every launch '
                               'of one eager move (tools/capture_traffic.py); mean per launch by kernel class', 'classes': agg}, indent=1) + '\n')
    print(dst.read_text())


if __name__ == '__main__':
    main()
