"""Dev helper: compact view of a bench.py JSON line on stdin."""
import json, sys
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        print(line[:300]); continue
    d = json.loads(line)
    print(f"value {d['value']:.4g} {d['unit']}  ms/step {d['ms_per_step']:.2f}  launches {d.get('gpu_launches')}  clocks {d.get('clocks')}")
    for k, v in (d.get("kernels") or {}).items():
        print(f"   {k:20s} {v['ms_per_step']:9.3f} ms  share {v['share_of_step']:.3f}  {v['achieved_gbs']} GB/s  frac {v['frac']}")
    print("   roofline", d.get("roofline"))
    print("   e2e", d.get("e2e"))
    print("   cpu", d.get("cpu_baseline"))
