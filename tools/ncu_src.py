"""Development tool: top stalled SASS instructions of an ncu report (source page), per kernel.
python tools/ncu_src.py rep [n] [kernel substring]"""
import csv, subprocess, sys
rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
want = sys.argv[3] if len(sys.argv) > 3 else ''
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
sections, seen = [], set()
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == 'Kernel Name':
        name, hdr = rows[i][1], rows[i + 1]
        j = i + 2
        while j < len(rows) and not (rows[j] and rows[j][0] == 'Kernel Name'):
            j += 1
        if name not in seen:   # the second copy of a kernel is the other source view
            seen.add(name)
            sections.append((name, hdr, rows[i + 2:j]))
        i = j
    else:
        i += 1
for name, hdr, body in sections:
    if want not in name:
        continue
    S = hdr.index('Warp Stall Sampling (All Samples)')
    I = hdr.index('Instructions Executed')
    stalls = [k for k, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    tot = sum(float(r[S] or 0) for r in body) or 1.
    print('kernel', name[:100], 'samples', tot, 'instructions', sum(float(r[I] or 0) for r in body))
    agg = {hdr[k]: sum(float(r[k] or 0) for r in body) for k in stalls}
    print('  stalls:', ' '.join('{}={:.0f}'.format(h[6:], v) for h, v in sorted(agg.items(), key=lambda t: -t[1]) if v))
    order = sorted(range(len(body)), key=lambda q: -float(body[q][S] or 0))[:n]
    for q in sorted(order):
        r = body[q]
        top = sorted(((float(r[k] or 0), hdr[k]) for k in stalls), reverse=True)[:2]
        print('{:5d} {:6.0f} {:5.1f}% x{:<8} {:70s} {}'.format(q, float(r[S] or 0), 100 * float(r[S] or 0) / tot, r[I], r[1].strip()[:70],
              ' '.join('{}={:.0f}'.format(h[6:], v) for v, h in top if v)))
