"""Per-launch table of ONE step from an ncu launch list (csv, --metrics gpu__time_duration.sum[,dram__bytes_*]):
    python tools/launch_table.py profiles/<launches>.csv [other.csv]   (two files: side-by-side durations)"""
import collections, csv, re, sys


def load(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    idx = {n: i for i, n in enumerate(rows[0])}
    L = collections.OrderedDict()
    for r in rows[1:]:
        d = L.setdefault(int(r[0]), {"name": r[idx["Kernel Name"]], "grid": r[idx["Grid Size"]]})
        d[r[idx["Metric Name"]]] = float(r[idx["Metric Value"]].replace(",", ""))
    ids = [i for i, d in L.items() if "patch_embed" in d["name"]]
    s, e = ids[-2], ids[-1]
    return [L[i] for i in range(s, e)]


def short(n):
    n = re.sub(r"\(.*", "", n).replace("void ", "").replace("<unnamed>::", "")
    return n[:44]


A = load(sys.argv[1])
B = load(sys.argv[2]) if len(sys.argv) > 2 else None
tot = 0.0
agg = collections.Counter()
for i, d in enumerate(A):
    t = d.get("gpu__time_duration.sum", 0) / 1e3
    tot += t
    agg[short(d["name"]).split("<")[0]] += t
    extra = ""
    if B and i < len(B):
        extra = f"   | {short(B[i]['name'])[:28]:28s} {B[i].get('gpu__time_duration.sum', 0) / 1e3:7.1f}"
    print(f"{i:3d} {short(d['name']):44s} {d['grid']:>14s} {t:7.1f} us  rd {d.get('dram__bytes_read.sum', 0) / 1e6:6.1f} wr {d.get('dram__bytes_write.sum', 0) / 1e6:6.1f} MB{extra}")
print(f"total {tot:.1f} us")
for k, v in agg.most_common():
    print(f"  {k:40s} {v:8.1f} us")
