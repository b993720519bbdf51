#!/usr/bin/env python3
"""ncu --metrics CSV of one serial-stage launch -> a small JSON summary (profiles/<tag>_rate_counters.json).
usage: rate_counters.py <ncu.csv> <out.json> [note]"""
import csv, json, sys
rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
r = list(csv.reader(rows)); h = r[0]; d = {}
for row in r[1:]:
    rec = dict(zip(h, row))
    d[rec["Metric Name"]] = float(rec["Metric Value"].replace(",", ""))
    kern = rec.get("Kernel Name", "")
cyc, inst = d["sm__cycles_active.sum"], d["smsp__inst_executed.sum"]
out = {"kernel": kern, "launch_ms": d["gpu__time_duration.sum"] / 1e6, "warp_instructions": inst,
       "issue_slots_busy": d["smsp__issue_active.sum"] / 4 / cyc, "ipc_per_sm": inst / cyc,
       "warps_resident_per_sm": d["smsp__warps_active.sum"] / cyc,
       "icc_hit_rate": d["sm__icc_requests_lookup_hit.sum"] / d["sm__icc_requests.sum"],
       "instructions_per_new_icache_line": inst / d["sm__icc_requests_lookup_miss_tag_miss.sum"],
       "stalled_warps_per_cycle": {k.replace("smsp__warps_issue_stalled_", "").replace(".sum", ""): round(v / cyc, 3)
                                   for k, v in d.items() if "stalled" in k}}
for k in ("l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum"):
    if k in d:
        out[k] = d[k]
if len(sys.argv) > 3:
    out["note"] = sys.argv[3]
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))
