"""Turn one `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` log of
`python bench.py ...` into the committed summaries: profiles/<tag>_launch_list_summary.md, <tag>_launches.csv and
profiles/ncu_traffic.json (dram bytes per step and launches per step of every kernel class; read by bench.py).

    python scripts/summarize_r02.py r02 gpurun_out/.../launches.csv <launches_per_step>
"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, path, per_step = sys.argv[1], sys.argv[2], int(sys.argv[3])
out_dir = os.path.join(ROOT, "profiles")

CLASS = [("radix_pass_kernel", "sort_radix_pass"), ("l2_scan_stream_kernel", "l2_scan_search"),
         ("l2_scan_bulk_kernel", "l2_scan_search"), ("l2_exact_kernel", "l2_exact_search"),
         ("exact_total_kernel", "l2_exact_search"), ("partition_kernel", "partition"),
         ("sort_prepare_kernel", "sort_prepare"), ("col_minmax_kernel", "sort_prepare"),
         ("col_range_kernel", "sort_prepare"), ("gather_payload_kernel", "gather_y_payload"),
         ("sort_fixup_kernel", "sort_fixup_ties"), ("finalize", "finalize_argbest"), ("mark_left", "mark_left_rows"),
         ("relabel_", "relabel_rows"), ("predict_kernel", "predict")]


def short(name):
    return name.replace("void <unnamed>::", "").replace("<unnamed>::", "").split("(")[0]


def to_num(v, unit):
    v = float(v.replace(",", ""))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1,
             "ms": 1e3, "msecond": 1e3}
    return v * scale.get(unit, 1)


rows = [r for r in csv.reader(open(path)) if len(r) > 5]
hdr = rows[0]
iid, ikn, imn, imu, imv = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
launches = collections.OrderedDict()
for r in rows[1:]:
    d = launches.setdefault(r[iid], {"kernel": short(r[ikn])})
    d[r[imn]] = to_num(r[imv], r[imu])
all_l = [d for d in launches.values() if "gpu__time_duration.sum" in d]
ours = [d for d in all_l if not d["kernel"].startswith(("at::", "void at::"))]
step = ours[-per_step:]                      # the last step's launches
per = collections.OrderedDict()
cls_bytes = collections.defaultdict(lambda: [0.0, 0])
for d in step:
    k = d["kernel"]
    e = per.setdefault(k, [0, 0.0, 0.0])
    e[0] += 1
    e[1] += d["gpu__time_duration.sum"]
    b = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    e[2] += b
    for prefix, c in CLASS:
        if k.startswith(prefix):
            cls_bytes[c][0] += b
            cls_bytes[c][1] += 1
            break
total = sum(e[1] for e in per.values())
with open(os.path.join(out_dir, f"{tag}_launch_list_summary.md"), "w") as f:
    f.write(f"# ncu launch list ({tag}) -- `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
            "dram__bytes_write.sum --clock-control none`\n\n")
    f.write(f"Command: `python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-inference` (10 M x 256, depth 10, "
            f"1 GPU); the {per_step} launches of the timed step (warm-up steps and torch's data-generation kernels "
            "dropped).  Times are cold-cache and serialised: compare SHARES with bench.py's "
            "`kernels[*].share_of_step`, not absolutes.  DRAM bytes = read + write of all launches of the kernel.\n\n")
    f.write("| kernel | launches | total ms | share | DRAM GB |\n|---|---|---|---|---|\n")
    for k, (n, us, b) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k}` | {n} | {us / 1e3:.2f} | {us / total:.3f} | {b / 1e9:.2f} |\n")
    f.write(f"\nTotal {total / 1e3:.1f} ms over {sum(e[0] for e in per.values())} launches.\n")
with open(os.path.join(out_dir, f"{tag}_launches.csv"), "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["index", "kernel", "duration_us", "dram_bytes"])
    for i, d in enumerate(step):
        w.writerow([i, d["kernel"], f"{d['gpu__time_duration.sum']:.3f}",
                    f"{d.get('dram__bytes_read.sum', 0.0) + d.get('dram__bytes_write.sum', 0.0):.0f}"])
tj = {c: {"bytes_per_step": v[0], "launches_per_step": v[1]} for c, v in cls_bytes.items()}
json.dump(tj, open(os.path.join(out_dir, "ncu_traffic.json"), "w"), indent=1)
print(open(os.path.join(out_dir, f"{tag}_launch_list_summary.md")).read())
