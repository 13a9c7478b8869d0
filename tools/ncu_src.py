"""Development tool: top stalled SASS instructions of an ncu report (source page).  python tools/ncu_src.py rep [n]"""
import csv, subprocess, sys
rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
body = rows[2:]
S = hdr.index('Warp Stall Sampling (All Samples)')
I = hdr.index('Instructions Executed')
stalls = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(float(r[S] or 0) for r in body)
print('kernel', rows[0][1][:100], 'samples', tot, 'instructions', sum(float(r[I] or 0) for r in body))
order = sorted(range(len(body)), key=lambda i: -float(body[i][S] or 0))[:n]
for i in sorted(order):
    r = body[i]
    top = sorted(((float(r[k] or 0), hdr[k]) for k in stalls), reverse=True)[:2]
    print('{:5d} {:6.0f} {:5.1f}% x{:<8} {:70s} {}'.format(i, float(r[S] or 0), 100 * float(r[S] or 0) / tot, r[I], r[1].strip()[:70],
          ' '.join('{}={:.0f}'.format(h[6:], v) for v, h in top if v)))
