import json, sys
for l in sys.stdin:
    l = l.strip()
    if not l.startswith("{"): continue
    d = json.loads(l)
    print(round(d["ms_per_step"], 3), "ms", {k: round(v, 3) for k, v in (d.get("phases_ms") or {}).items()}, "launches", d.get("gpu_launches"),
          "roof", (d.get("roofline") or {}).get("frac"))
