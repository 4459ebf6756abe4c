"""Top SASS instructions by warp-stall samples from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.
usage: ncu_hot_sass.py cs.csv [file_substr] [line_lo] [line_hi] [top_n]"""
import csv, sys
path = sys.argv[1]
fsub = sys.argv[2] if len(sys.argv) > 2 else ""
lo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4]) if len(sys.argv) > 4 else 10**9
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
rows = list(csv.reader(open(path)))
cur = None; hdr = None; line = None; items = []
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) >= 2 and r[0] == 'Function Name': continue
    if len(r) > 3 and r[0] == 'Line No': hdr = r; si = hdr.index('# Samples'); continue
    if not hdr or len(r) <= si: continue
    if r[0].isdigit(): line = int(r[0]); continue
    if r[0] == '' and r[2].startswith('0x'):
        try: n = int(r[si])
        except ValueError: continue
        items.append((n, cur, line, r[3].strip(), r))
names = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
idx = {h: hdr.index(h) for h in names}
sel = [it for it in items if fsub in it[1] and lo <= (it[2] or 0) < hi]
tot = sum(it[0] for it in items)
print('samples in selection', sum(it[0] for it in sel), 'of', tot)
for n, f, l, sass, r in sorted(sel, key=lambda x: -x[0])[:top]:
    st = sorted(((int(r[idx[h]] or 0), h[6:]) for h in names), reverse=True)[:2]
    print(f"{n:6d} {f}:{l:4d} {sass[:64]:64s} {st}")
