#!/usr/bin/env python
"""Turn ncu outputs from gpurun_out/ into the small text summaries committed under profiles/.

  python tools/ncu_summary.py launches gpurun_out/r1_launches.csv profiles/r1_launches.md
  python tools/ncu_summary.py full gpurun_out/r1_scan_full.ncu-rep profiles/r1_scan_full.md [profiles/scan_traffic.json]
"""
import csv
import io
import json
import re
import subprocess
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"<unnamed>::", "", name)
    m = re.match(r"(?:void )?([\w:]+(?:<[^(]{0,40}>)?)", name)
    s = m.group(1) if m else name
    return s[:90]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if r and r[0].isdigit()]
    recs = [(int(r[0]), short(r[4]), r[7], r[8], float(r[-1]) / 1e6) for r in rows]   # ms
    agg = OrderedDict()
    for _, k, _, _, ms in recs:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(v[1] for v in agg.values())
    # the query step: from one itq_hash launch (Q rows) to the next
    scan_ids = [i for i, (_, k, _, _, _) in enumerate(recs) if k.startswith("hamming_scan_kernel")]
    out = ["# ncu launch list summary (`%s`)" % src, "",
           "`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: shares, not absolutes).",
           "", "%d launches, %.1f ms of kernel time in total." % (len(recs), tot), "",
           "| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
    for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
        out.append("| `%s` | %d | %.3f | %.1f%% |" % (k, c, ms, 100 * ms / tot))
    # one query step = launches between two consecutive batched scans (grid.y > 1)
    big = [i for i in scan_ids if not recs[i][3].endswith(", 1, 1)")]
    if len(big) >= 3:
        a, b = big[-3], big[-2]
        step = recs[a:b]
        st = sum(r[4] for r in step)
        out += ["", "## One query step (launch %d .. %d, between two batched scans)" % (recs[a][0], recs[b][0] - 1), "",
                "| kernel | grid | block | ms | share of step |", "|---|---|---|---:|---:|"]
        for _, k, blk, grd, ms in step:
            out.append("| `%s` | %s | %s | %.4f | %.2f%% |" % (k, grd, blk, ms, 100 * ms / st))
        out.append("")
        out.append("step total %.3f ms; `ham_filter_*` (tensor-core scan) share %.2f%%; `hamming_scan_kernel` share %.2f%%" %
                   (st, 100 * sum(r[4] for r in step if r[1].startswith("ham_filter")) / st,
                    100 * sum(r[4] for r in step if r[1].startswith("hamming_scan_kernel")) / st))
    open(dst, "w").write("\n".join(out) + "\n")
    print("wrote", dst)


KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
    "smsp__cycles_active.avg", "sm__cycles_active.avg",
]


def full(src, dst, traffic_json=None):
    txt = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    out = ["# ncu --set full summary (`%s`)" % src, ""]
    traffic = None
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        out += ["## `%s`  grid %s block %s" % (short(d.get("Kernel Name", "?")), d.get("Grid Size"), d.get("Block Size")), "",
                "| metric | value | unit |", "|---|---:|---|"]
        for k in KEYS:
            if k in d and d[k] != "":
                out.append("| %s | %s | %s |" % (k, d[k], u.get(k, "")))
        stalls = sorted(((float(d[k]), k) for k in hdr if k.startswith("smsp__average_warp") and "issue_stalled" in k
                         and k.endswith("_per_issue_active.ratio") and d[k] not in ("", "n/a")), reverse=True)[:8]
        if stalls:
            out += ["", "top warp-stall reasons (warps per issue-active cycle):", ""]
            out += ["* %s = %.3f" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v)
                    for v, k in stalls]
        out.append("")

        def to_bytes(key):
            v = float(d[key])
            un = u.get(key, "byte").lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(un, 1)
        try:
            traffic = {"kernel": short(d["Kernel Name"]), "dram_bytes_per_launch": to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"),
                       "dram_read_bytes": to_bytes("dram__bytes_read.sum"), "dram_write_bytes": to_bytes("dram__bytes_write.sum"),
                       "gpu_time_ms_under_ncu": float(d["gpu__time_duration.sum"]) * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(u.get("gpu__time_duration.sum", "ns"), 1e-6),
                       "source": src}
        except Exception as e:
            print("traffic:", e)
    open(dst, "w").write("\n".join(out) + "\n")
    print("wrote", dst)
    if traffic_json and traffic:
        json.dump(traffic, open(traffic_json, "w"), indent=1)
        print("wrote", traffic_json)


def traffic(src, dst, traffic_json, title, kernel_prefix):
    """ncu --csv launch list with several metrics per launch (time + DRAM bytes) -> table + per-step traffic json."""
    rows = list(csv.reader(open(src)))
    hdr = next(r for r in rows if "Metric Name" in r)
    ci = {h: i for i, h in enumerate(hdr)}
    per = OrderedDict()
    for r in rows:
        if len(r) != len(hdr) or not r[0].isdigit():
            continue
        d = per.setdefault(int(r[ci["ID"]]), {"kernel": short(r[ci["Kernel Name"]])})
        v = float(r[ci["Metric Value"]].replace(",", ""))
        unit = r[ci["Metric Unit"]].lower()
        scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6,
                 "nsecond": 1e-3, "usecond": 1, "msecond": 1e3, "second": 1e6}.get(unit, 1)
        d[r[ci["Metric Name"]]] = v * scale
    out = ["# %s" % title, "", "`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none` "
           "(cold-cache, serialised: shares matter, not absolutes)", "", "| # | kernel | time (us) | DRAM read (MB) | DRAM write (MB) |",
           "|---|---|---:|---:|---:|"]
    tot = 0.0
    n_k = 0
    for i, (_, d) in enumerate(per.items()):
        rd, wr = d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
        out.append("| %d | `%s` | %.1f | %.2f | %.2f |" % (i, d["kernel"], d.get("gpu__time_duration.sum", 0.0), rd / 1e6, wr / 1e6))
        if d["kernel"].startswith(kernel_prefix):
            tot += rd + wr
            n_k += 1
    out += ["", "`%s`: %d launches, %.1f MB of DRAM traffic in total." % (kernel_prefix, n_k, tot / 1e6)]
    open(dst, "w").write("\n".join(out) + "\n")
    json.dump({"kernel": kernel_prefix, "source": "%s (ncu dram__bytes_read.sum + dram__bytes_write.sum, summed over the "
               "launches of one batch)" % dst, "dram_bytes_per_step": tot, "launches_per_step": n_k, "workload": title},
              open(traffic_json, "w"), indent=1)
    print("wrote", dst, traffic_json)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == "traffic":
        traffic(*sys.argv[2:7])
    else:
        full(*sys.argv[2:])
