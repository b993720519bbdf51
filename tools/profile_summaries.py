#!/usr/bin/env python3
"""Turn the ncu outputs a gpurun call left in gpurun_out/ into the text summaries committed under profiles/.
usage: profile_summaries.py <tag>   (expects gpurun_out/<tag>_launches.csv, <tag>_rate.ncu-rep[, <tag>_pack.ncu-rep])"""
import collections, csv, os, subprocess, sys
tag = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

def launches():
    src = os.path.join(G, tag + "_launches.csv")
    if not os.path.exists(src):
        return
    rows = list(csv.reader(open(src)))
    i = [k for k, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[i]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[i + 1:]:
        if len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        u = d["Metric Unit"]
        ms = v / 1e6 if u.startswith("n") else (v / 1e3 if u.startswith("u") else v)
        name = d["Kernel Name"].split("(")[0]
        agg[name][0] += 1
        agg[name][1] += ms
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(P, tag + "_launch_summary.txt"), "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 1 --warmup 1 "
                "--no-cpu-baseline\n(first 400 launches; serialised and cold-cache under the "
                "profiler: compare SHARES, not absolutes)\n\n")
        f.write("%-22s %6s %12s %7s\n" % ("kernel", "count", "total ms", "share"))
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-22s %6d %12.3f %6.1f%%\n" % (k, v[0], v[1], 100 * v[1] / tot))
    subprocess.run(["cp", src, os.path.join(P, tag + "_launches.csv")])

def details(name, cmdline):
    rep = os.path.join(G, "%s_%s.ncu-rep" % (tag, name))
    if not os.path.exists(rep):
        return
    out = subprocess.run(["ncu", "-i", rep, "--page", "details"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    d = dict(zip(rows[0], rows[2]))
    st = {k: float(v.replace(",", "")) for k, v in d.items() if k.startswith("smsp__pcsamp_warps_issue_stalled_")
          and not k.endswith("_not_issued") and v not in ("", "n/a")}
    tot = sum(st.values()) or 1
    with open(os.path.join(P, "%s_%s_ncu_details.txt" % (tag, name)), "w") as f:
        f.write(cmdline + "\n\nwarp stall samples (all):\n")
        for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:10]:
            f.write("   %-36s %6.2f%%\n" % (k[len("smsp__pcsamp_warps_issue_stalled_"):], 100 * v / tot))
        f.write("\ndram__bytes_read.sum %s GB, dram__bytes_write.sum %s GB, smsp__inst_executed.sum %s\n\n" %
                (d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum"), d.get("smsp__inst_executed.sum")))
        f.write(out)
    if name == "rate":
        fn = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_by_function.py"), rep,
                             os.path.join(ROOT, "hmp3_b200", "_lib", "libhmp3_b200.so"), "k_rate"],
                            stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
        open(os.path.join(P, "%s_rate_by_function.txt" % tag), "w").write(cmdline + "\n\n" + fn)

launches()
details("rate", "ncu --set full --clock-control none --import-source on -k regex:k_rate$ -s 2 -c 1 python tools/quick_bench.py 4736 6   (4736 streams x 6 s, 128 granules per launch)")
details("others", "ncu --set full --clock-control none --import-source on -k regex:'k_pack|k_polyphase|k_hybrid|k_psy_stage1|k_prepare$|k_psy_stage2' -s 6 -c 6 python tools/quick_bench.py 4736 6   (one launch of each Phase A / packing kernel)")
print(open(os.path.join(P, tag + "_launch_summary.txt")).read())
