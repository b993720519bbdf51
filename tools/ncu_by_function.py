#!/usr/bin/env python3
"""Aggregate an ncu source-page capture of a kernel per (noinline) device function.

usage: ncu_by_function.py <report.ncu-rep> <library.so> <kernel-substring> [units]
(units = granules x streams of the captured launch: adds per-function code footprint columns -- static instructions,
instructions executed at least once per two units ("hot"), dynamic instructions per unit)
Joins `ncu --page source --print-source sass` (per-instruction samples / executed counts) with the
function symbol table of the cubin extracted from the library."""
import csv, subprocess, sys, os, re, tempfile, collections

rep, so, kern = sys.argv[1], sys.argv[2], sys.argv[3]
units = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
syms = []
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    out = subprocess.run(["readelf", "-sW", os.path.join(tmp, f)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    for line in out.splitlines():
        p = line.split()
        if len(p) >= 8 and p[3] == "FUNC" and ("$" + "_Z") in p[7] and re.search(r"\d" + kern + "E", p[7].split("$")[1]):
            name = p[7].split("$")[2]
            m = re.search(r"hmp3(\d+)([A-Za-z_0-9]+)", name)
            short = m.group(2)[:int(m.group(1))] if m else name
            syms.append((int(p[1], 16), int(p[2]), short))
syms.sort()
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE,
                     stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = None
base = None
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0, 0, 0, 0])
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        continue
    if not hdr or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    a = int(d["Address"], 16)
    if base is None:
        base = a
    off = a - base
    fn = "(kernel body)"
    for s0, sz, nm in syms:
        if s0 <= off < s0 + sz:
            fn = nm
            break
    agg[fn][0] += int(d["# Samples"] or 0)
    agg[fn][1] += int(d["Instructions Executed"] or 0)
    agg[fn][2] += int(d["Thread Instructions Executed"] or 0)
    agg[fn][3] += int(d.get("stall_no_inst") or 0)
    agg[fn][4] += int(d.get("stall_long_sb") or 0)
    agg[fn][5] += int(d.get("stall_wait") or 0)
    agg[fn][6] += 1
    if units and int(d["Instructions Executed"] or 0) >= 0.5 * units:
        agg[fn][7] += 1
ts = sum(v[0] for v in agg.values()) or 1
ti = sum(v[1] for v in agg.values()) or 1
print("%-28s %8s %8s %14s %8s %8s %8s %8s" % ("function", "samples%", "instr%", "warp-instr", "thr/inst", "no_inst%", "long_sb%", "wait%"))
for fn, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%-28s %7.2f%% %7.2f%% %14d %8.1f %7.2f%% %7.2f%% %7.2f%%" % (fn, 100 * v[0] / ts, 100 * v[1] / ti, v[1], v[2] / max(v[1], 1),
                                                            100 * v[3] / ts, 100 * v[4] / ts, 100 * v[5] / ts))
print("total warp instructions", ti, "samples", ts)
if units:
    print()
    print("%-28s %8s %8s %8s %10s %8s" % ("function", "static", "hot", "hot KB", "dyn/unit", "samples%"))
    for fn, v in sorted(agg.items(), key=lambda kv: -kv[1][7]):
        print("%-28s %8d %8d %8.1f %10.0f %7.2f%%" % (fn, v[6], v[7], v[7] * 16 / 1024, v[1] / units, 100 * v[0] / ts))
    print("hot footprint: %.1f KB of %.1f KB" % (sum(v[7] for v in agg.values()) * 16 / 1024, sum(v[6] for v in agg.values()) * 16 / 1024))
