"""Development tool: executed warp-instructions by opcode from an ncu report's source page, per kernel.
python tools/ncu_ops.py rep [kernel substring] [full]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ''
full = len(sys.argv) > 3
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
seen = set()
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == 'Kernel Name':
        name, hdr = rows[i][1], rows[i + 1]
        j = i + 2
        while j < len(rows) and not (rows[j] and rows[j][0] == 'Kernel Name'):
            j += 1
        if name not in seen and want in name:
            seen.add(name)
            I = hdr.index('Instructions Executed')
            ops = collections.Counter()
            for r in rows[i + 2:j]:
                t = r[1].strip().split()
                if not t or len(r) <= I:
                    continue
                op = t[1] if t[0].startswith('@') else t[0]
                ops[op if full else op.split('.')[0]] += float(r[I] or 0)
            tot = sum(ops.values()) or 1.
            print('kernel', name[:100], 'total', tot)
            for op, v in ops.most_common(32):
                print('  {:14s} {:12.0f} {:5.1f}%'.format(op, v, 100 * v / tot))
        i = j
    else:
        i += 1
