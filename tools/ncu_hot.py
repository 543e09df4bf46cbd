"""Hot spots of one kernel from an `ncu --page source --csv` dump: SASS lines sorted by stall samples, with the dominant stall reason.
    ncu -i rep.ncu-rep --page source --csv --kernel-name regex:<k> --launch-skip i --launch-count 1 > src.csv; python tools/ncu_hot.py src.csv [N] [--seq lo hi]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr) and r[0].startswith("0x")]       # SASS lines only
n = int(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else 40
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[idx["# Samples"]] or 0) for r in data)
inst = sum(int(r[idx["Instructions Executed"]] or 0) for r in data)
print("total samples %d, warp instructions executed %d, SASS lines %d" % (tot, inst, len(data)))
agg = {s: sum(int(r[idx[s]] or 0) for r in data) for s in stalls}
print("stalls: " + " ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
ops = {}
for r in data:
    op = r[idx["Source"]].split()[0] if r[idx["Source"]].split() else "?"
    if op.startswith("@"):
        op = r[idx["Source"]].split()[1]
    op = op.split(".")[0]
    ops[op] = ops.get(op, 0) + int(r[idx["Instructions Executed"]] or 0)
print("executed mix: " + " ".join("%s %.1f%%" % (k, 100.0 * v / max(inst, 1)) for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:14]))
if "--seq" in sys.argv:
    lo, hi = int(sys.argv[sys.argv.index("--seq") + 1]), int(sys.argv[sys.argv.index("--seq") + 2])
    for i, r in enumerate(data[lo:hi]):
        s = int(r[idx["# Samples"]] or 0)
        top = max(stalls, key=lambda k: int(r[idx[k]] or 0))
        print("%5d %6d %-10s %s" % (lo + i, s, top[6:] if s else "", r[idx["Source"]].strip()[:110]))
else:
    order = sorted(range(len(data)), key=lambda i: -int(data[i][idx["# Samples"]] or 0))[:n]
    for i in order:
        r = data[i]
        s = int(r[idx["# Samples"]] or 0)
        top = max(stalls, key=lambda k: int(r[idx[k]] or 0))
        print("%5d %6d %4.1f%% %-10s %s" % (i, s, 100.0 * s / max(tot, 1), top[6:], r[idx["Source"]].strip()[:110]))
