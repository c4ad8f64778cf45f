"""Summarise an `ncu --page source --csv` dump: hottest SASS instructions with their dominant stall reasons.

    ncu -i prof.ncu-rep --page source --csv --kernel-name regex:mlp_tc > src.csv ; python scripts/ncu_hot.py src.csv [topN]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
for b in blocks[:1]:
    hdr, data = b["hdr"], b["data"]
    idx = {h: i for i, h in enumerate(hdr)}
    tot = sum(int(r[idx["# Samples"]]) for r in data)
    print(b["name"][:100], "total samples", tot, "instructions", len(data))
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {}
    for r in data:
        for h in stalls:
            agg[h] = agg.get(h, 0) + int(r[idx[h]])
    print("stall totals:", sorted(((v, k[6:]) for k, v in agg.items()), reverse=True)[:8])
    top = sorted(range(len(data)), key=lambda i: -int(data[i][idx["# Samples"]]))[:topn]
    for i in sorted(top):
        r = data[i]
        s = int(r[idx["# Samples"]])
        st = sorted(((int(r[idx[h]]), h[6:]) for h in stalls), reverse=True)[:2]
        print(f"{i:5d} {r[idx['Source']].strip()[:66]:66s} {s:6d} {100 * s / tot:5.1f}% {st} exec={r[idx['Instructions Executed']]}")
