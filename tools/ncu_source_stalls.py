"""Summarise the source page of an ncu report: per-SASS-instruction stall samples over an address range.
Usage: ncu -i rep --page source --csv --launch-skip K --launch-count 1 > src.csv
       python tools/ncu_source_stalls.py src.csv [lo_hex hi_hex] [--top N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
lo = int(sys.argv[2], 16) if len(sys.argv) > 3 and not sys.argv[2].startswith("--") else 0
hi = int(sys.argv[3], 16) if len(sys.argv) > 3 and not sys.argv[2].startswith("--") else 1 << 62
base = None
tot = {}
nsamp = 0
lines = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    a = int(r[idx["Address"]], 16) if r[idx["Address"]].startswith("0x") or all(c in "0123456789abcdef" for c in r[idx["Address"]].lower()) else None
    if a is None:
        continue
    if base is None:
        base = a
    off = a - base
    if not (lo <= off < hi):
        continue
    n = int(r[idx["# Samples"]] or 0)
    st = {s[6:]: int(r[idx[s]] or 0) for s in stalls if int(r[idx[s]] or 0)}
    for k, v in st.items():
        tot[k] = tot.get(k, 0) + v
    nsamp += n
    lines.append((off, r[idx["Source"]], n, int(r[idx["Instructions Executed"]] or 0), st,
                  r[idx["L1 Wavefronts Shared"]], r[idx["L1 Wavefronts Shared Ideal"]]))
if "--list" in sys.argv:
    for off, src, n, ex, st, wf, wfi in lines:
        print("%05x %-62s %5d %8d %s %s" % (off, src[:62], n, ex, sorted(st.items(), key=lambda kv: -kv[1])[:3],
                                              ("wf %s/%s" % (wf, wfi)) if wf not in ("", "0") else ""))
print("range %x..%x: %d instructions, %d samples; stall totals:" % (lo, hi if hi < 1 << 62 else 0, len(lines), nsamp),
      sorted(tot.items(), key=lambda kv: -kv[1]))
