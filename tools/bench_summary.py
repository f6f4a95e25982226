"""Prints one line per bench JSON file: value, ms/step, roofline fraction (and the `also` entries)."""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as e:
        print(path, "unreadable:", e)
        continue
    r = d.get("roofline", {})
    print(f"{path}: {d.get('value', 0):.1f} {d.get('unit')} | {d.get('ms_per_step', 0):.4f} ms/step | frac {r.get('frac', 0):.3f} | e2e {d.get('e2e', {}).get('value')} | {d.get('path') or d.get('config', {}).get('path')} | base {d.get('strong_scaling_base', {}).get('value')}")
    for a in d.get("also", []):
        print(f"    also {a.get('workload', '')[:60]}: {a.get('value', 0):.1f} | {a.get('ms_per_step', 0):.4f} ms | frac {a.get('roofline', {}).get('frac', 0):.3f} {a.get('unavailable', '')}")
