"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line.
usage: python scratch/ncu_lines.py export.csv [top_n] [file_filter]"""
import csv, collections, sys
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40; filt = sys.argv[3] if len(sys.argv) > 3 else None
hdr = None; cur = None; line = None; src = None
agg = collections.defaultdict(lambda: [0, 0, ""]); stall = collections.defaultdict(collections.Counter)
seen_func = collections.Counter()
for r in csv.reader(open(path)):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": seen_func[(cur, r[1])] += 1; inst = seen_func[(cur, r[1])]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None: continue
    if r[0].isdigit():
        line = int(r[0]); src = r[1][:95]; continue
    if r[0] == "" and len(r) > 7 and r[2].startswith("0x"):
        try: s = int(r[6] or 0); ins = int(r[7] or 0)
        except ValueError: continue
        k = (cur, line); agg[k][0] += s; agg[k][1] += ins; agg[k][2] = src
        for i in range(32, 49):
            try: stall[k][hdr[i]] += int(r[i] or 0)
            except ValueError: pass
items = [(v[0], k[0], k[1], v[1], v[2]) for k, v in agg.items() if not filt or filt in (k[0] or "")]
print("total samples", sum(i[0] for i in items), "total warp-instructions", sum(i[3] for i in items))
byf = collections.Counter(); byi = collections.Counter()
for i in items: byf[i[1]] += i[0]; byi[i[1]] += i[3]
print("samples by file", dict(byf)); print("instr by file", dict(byi))
for it in sorted(items, reverse=True)[:topn]:
    print("%5d %-16s %4d %9d  %s  %s" % (it[0], it[1], it[2], it[3], it[4], stall[(it[1], it[2])].most_common(2)))
