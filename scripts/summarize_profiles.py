"""Dev helper: turn gpurun_out ncu artefacts into the committed summaries under profiles/."""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]            # e.g. r01
launch_csv = sys.argv[2]     # ncu --metrics gpu__time_duration.sum --csv log
rep = sys.argv[3]            # .ncu-rep of the --set full capture
out_dir = os.path.join(ROOT, "profiles")

def short(name):
    name = name.replace("void <unnamed>::", "").replace("<unnamed>::", "")
    return name.split("(")[0]

# ---- launch list -------------------------------------------------------------------------------
rows = [r for r in csv.reader(open(launch_csv)) if len(r) > 5]
hdr = rows[0]
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
mu = hdr.index("Metric Unit")
per = collections.OrderedDict()
lines = []
for r in rows[1:]:
    try:
        v = float(r[mv].replace(",", ""))
    except ValueError:
        continue
    unit = r[mu]
    us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
    k = short(r[kn])
    d = per.setdefault(k, [0, 0.0])
    d[0] += 1
    d[1] += us
    lines.append((k, us))
total = sum(d[1] for d in per.values())
with open(os.path.join(out_dir, f"{tag}_launch_list_summary.md"), "w") as f:
    f.write(f"# ncu launch list ({tag}) -- `ncu --metrics gpu__time_duration.sum --clock-control none`\n\n")
    f.write("Command: `python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e` (10 M x 256, depth 10, 1 GPU), "
            "the 2 timed steps (launches of the 3 warm-up steps skipped).  Times are cold-cache and serialised: "
            "compare SHARES with bench.py's `kernels[*].share_of_step`, not absolutes.\n\n")
    f.write("| kernel | launches | total ms | share |\n|---|---|---|---|\n")
    for k, (n, us) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k}` | {n} | {us/1e3:.2f} | {us/total:.3f} |\n")
    f.write(f"\nTotal {total/1e3:.1f} ms over {sum(d[0] for d in per.values())} launches "
            f"({total/2e3:.1f} ms per step).\n")
with open(os.path.join(out_dir, f"{tag}_launches.csv"), "w") as f:
    f.write("index,kernel,duration_us\n")
    for i, (k, us) in enumerate(lines):
        f.write(f"{i},{k},{us:.3f}\n")

# ---- full capture ------------------------------------------------------------------------------
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
want = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("smsp__inst_executed.sum", "warp inst"), ("lts__t_sector_hit_rate.pct", "L2 hit %")]
kn = hdr.index("Kernel Name")
def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)
traffic = collections.defaultdict(list)
with open(os.path.join(out_dir, f"{tag}_ncu_full_summary.md"), "w") as f:
    f.write(f"# ncu --set full summary ({tag})\n\n`ncu --set full --clock-control none --import-source on` on "
            "`python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e` (10 M x 256), launches after the warm-up.\n\n")
    f.write("| kernel | " + " | ".join(n for _, n in want) + " |\n|---|" + "---|" * len(want) + "\n")
    for r in data:
        cells = []
        for m, _ in want:
            if m in hdr:
                i = hdr.index(m)
                v = r[i]
                try:
                    v = f"{float(v.replace(',', '')):.4g}"
                except ValueError:
                    pass
                cells.append(f"{v} {units[i]}".strip())
            else:
                cells.append("-")
        name = r[kn].replace("void <unnamed>::", "").replace("<unnamed>::", "")
        name = name.split("(")[0]
        f.write(f"| `{name}` | " + " | ".join(cells) + " |\n")
        rd = to_bytes(r[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")])
        wr = to_bytes(r[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
        traffic[name].append(rd + wr)
cls = {"radix_pass_kernel<unsigned int, 0, 0>": "sort_radix_pass", "l2_scan_stream_kernel": "l2_scan_search",
       "partition_kernel": "partition", "sort_prepare_kernel": "sort_prepare",
       "gather_payload_kernel": "gather_y_payload", "sort_fixup_kernel": "sort_fixup_ties"}
tj = {}
for name, vals in traffic.items():
    for k, c in cls.items():
        if name.startswith(k):
            tj[c] = sum(vals) / len(vals)
json.dump(tj, open(os.path.join(out_dir, "ncu_traffic.json"), "w"), indent=1)
print(open(os.path.join(out_dir, f"{tag}_launch_list_summary.md")).read())
print(json.dumps(tj, indent=1))
