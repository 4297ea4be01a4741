"""Dev helper: summarise the per-line warp-stall samples of one kernel from an .ncu-rep."""
import csv, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}",
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
data = rows[2:]
isamp = hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter()
lines = []
for r in data:
    try:
        n = int(r[isamp])
    except Exception:
        continue
    st = {h: int(r[i] or 0) for i, h in stall_cols}
    for h, v in st.items():
        tot[h] += v
    lines.append((n, r[1], st))
total = sum(n for n, _, _ in lines)
print("total samples", total)
print("by reason:", ", ".join(f"{h[6:]}={v*100//max(total,1)}%" for h, v in tot.most_common(8)))
for n, src, st in sorted(lines, key=lambda x: -x[0])[:int(sys.argv[4]) if len(sys.argv) > 4 else 14]:
    top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"{n*100/max(total,1):5.1f}%  {src[:110]:110s} {top[0][0][6:]}:{top[0][1]} {top[1][0][6:]}:{top[1][1]}")
