"""Top stall sites of a kernel from an .ncu-rep captured with --import-source on:
python tools/ncu_sass_top.py report.ncu-rep [N]"""
import csv, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '--print-source', 'sass'],
                     capture_output=True, text=True).stdout.splitlines()
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(out))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[hi]
body = [dict(zip(hdr, r)) for r in rows[hi + 1:] if len(r) == len(hdr)]
tot = sum(int(b['# Samples'] or 0) for b in body)
insts = sum(int(b['Instructions Executed'] or 0) for b in body)
print('total samples', tot, 'warp instructions', insts, 'sass lines', len(body))
# opcode histogram (executed)
import collections
hist = collections.Counter()
for b in body:
    op = b['Source'].split()[0] if not b['Source'].startswith('@') else b['Source'].split()[1]
    hist[op.split('.')[0]] += int(b['Instructions Executed'] or 0)
print('opcode mix:', ', '.join('%s %.1f%%' % (k, 100.0 * v / insts) for k, v in hist.most_common(18)))
keys = ['stall_long_sb', 'stall_short_sb', 'stall_wait', 'stall_barrier', 'stall_math', 'stall_mio', 'stall_lg', 'stall_not_selected']
for idx, b in enumerate(body):
    b['_i'] = idx
top = sorted(body, key=lambda b: -int(b['# Samples'] or 0))[:n]
for b in sorted(top, key=lambda b: b['_i']):
    st = ' '.join('%s=%s' % (k[6:], b[k]) for k in keys if b.get(k) not in (None, '', '0'))
    print('%5d %5.1f%% %-70s %s' % (b['_i'], 100.0 * int(b['# Samples']) / tot, b['Source'][:70], st))
