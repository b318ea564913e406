#!/usr/bin/env python
"""gpurun_out/parity_margins.jsonl (written by the -m gpu tests through tests/helpers.record_margin) ->
profiles/r02_parity_margins.json: every parity check's ACHIEVED error next to its tolerance.
    python tools/collect_margins.py [in.jsonl] [out.json]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "parity_margins.jsonl")
dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "r02_parity_margins.json")
rows = {}
with open(src) as fh:
    for line in fh:
        r = json.loads(line)
        rows[(r["case"], r["what"])] = r          # the last run of a check wins
out = sorted(rows.values(), key=lambda r: (r["case"], r["what"]))
for r in out:
    r["fraction_of_tolerance"] = r["value"] / r["tol"] if r["tol"] else None
with open(dst, "w") as fh:
    json.dump({"note": "value = max|ours - ref| / max|ref| unless the check says otherwise; rms_rel = RMS(diff) / RMS(ref)",
               "checks": out}, fh, indent=1)
print(f"{len(out)} checks -> {dst}; worst fraction of tolerance: "
      f"{max((r['fraction_of_tolerance'] or 0) for r in out):.3f}")
