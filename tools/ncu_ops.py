"""Development tool: executed warp-instructions by opcode from an ncu report's source page.  python tools/ncu_ops.py rep"""
import csv, subprocess, sys, collections
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; body = rows[2:]
I = hdr.index('Instructions Executed')
ops = collections.Counter()
for r in body:
    t = r[1].strip().split()
    if not t: continue
    op = t[1] if t[0].startswith('@') else t[0]
    ops[op if len(sys.argv) > 2 else op.split(".")[0]] += float(r[I] or 0)
tot = sum(ops.values())
print('total', tot)
for op, v in ops.most_common(40):
    print('{:12s} {:12.0f} {:5.1f}%'.format(op, v, 100 * v / tot))
