"""Per-source-line summary of an ncu source page dumped with
   ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > X.csv
Prints, for the first profiled launch: instructions executed, average active threads, stall samples per line."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thresh = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
cur = None; out = []; seen = set(); launches = 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split('/')[-1]
        if cur in seen: break
        seen.add(cur)
        continue
    if len(r) > 12 and r[0].isdigit():
        try:
            ln = int(r[0]); inst = int(r[7]); thr = int(r[8]); samples = int(r[6])
        except ValueError:
            continue
        if inst > 0: out.append((cur, ln, inst, thr / inst, samples, r[1][:100]))
tot = sum(o[2] for o in out); ts = sum(o[4] for o in out)
print("total warp inst", tot, "samples", ts, "avg threads", sum(o[2]*o[3] for o in out)/tot)
for o in sorted(out, key=lambda x: (x[0], x[1])):
    if o[2] > tot * thresh or o[4] > ts * thresh:
        print("%-14s %4d inst %5.1f%% avgthr %5.1f samp %5.1f%% | %s" % (o[0], o[1], 100 * o[2] / tot, o[3], 100 * o[4] / ts, o[5]))
