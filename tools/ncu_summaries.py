"""Turn the ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

  python tools/ncu_summaries.py shares  gpurun_out/launches_X.csv  profiles/launch_shares_X.csv  "<command>"
  python tools/ncu_summaries.py metrics gpurun_out/Y.ncu-rep       profiles/Y_ncu.txt            "<command>"
  python tools/ncu_summaries.py side    gpurun_out/expert_ffn.ncu-rep profiles/expert_ffn_ncu_latest.json "<command>"
      (the roofline side-data bench.py attaches to its line: DRAM traffic + tensor-pipe utilisation of the two grouped
       expert GEMMs, tagged with the commit the capture was taken at)
"""
import csv, gzip, io, re, subprocess, sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max.per_second",
    "launch__block_size", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def shares(src, dst, cmd):
    rows = [l for l in open(src, newline="") if l.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(rows)))
    agg = {}
    for r in rd:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*$", "", r["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "").replace(
            "(anonymous namespace)::", "").replace(",", ";").strip()
        v = float(r["Metric Value"].replace(",", ""))
        us = v / 1e3 if r["Metric Unit"] == "ns" else (v if r["Metric Unit"] in ("us", "usecond") else v * 1e3)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write("# %s\n# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes\n" % cmd)
        f.write("kernel,launches,total_us,share_pct,avg_us\n")
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%s,%d,%.1f,%.2f,%.2f\n" % (k, n, us, 100 * us / tot, us / n))
    with gzip.open(dst.replace("launch_shares", "launches").replace(".csv", "_ncu.csv.gz"), "wt") as g:
        g.writelines(rows)


def metrics(src, dst, cmd):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    head, units, body = rd[0], rd[1], rd[2:]
    kcol = head.index("Kernel Name")
    with open(dst, "w") as f:
        f.write("# %s\n" % cmd)
        for row in body:
            f.write("kernel: %s\n" % row[kcol])
            for i, h in enumerate(head):
                short = h.split(".", 2)[-1] if h.count(".") >= 2 and h.split(".")[0].isupper() else h
                if (short in KEEP or h in KEEP or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")
                                                 and "not_issued" not in h)) and row[i] not in ("", "0"):
                    f.write("  %-78s %s %s\n" % (short, row[i], units[i]))


def side(src, dst, cmd):
    import json
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    head, body = rd[0], rd[2:]
    col = lambda name: next(i for i, h in enumerate(head) if h.endswith(name))
    num = lambda v: float(v.replace(",", ""))
    rows = []
    for row in body[:2]:                      # the up- and the down-projection launch of one expert FFN
        rows.append({"kernel": row[head.index("Kernel Name")][:60],
                     "dram_bytes": num(row[col("dram__bytes_read.sum")]) + num(row[col("dram__bytes_write.sum")]),
                     "dram_unit": rd[1][col("dram__bytes_read.sum")],
                     "tensor_pipe_pct": num(row[col("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")]),
                     "duration": num(row[col("gpu__time_duration.sum")]), "duration_unit": rd[1][col("gpu__time_duration.sum")]})
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    json.dump({"commit": commit, "source": dst.replace(".json", ".txt") + " (ncu --set full, " + cmd + ")",
               "traffic": sum(r["dram_bytes"] * scale.get(r["dram_unit"], 1.0) for r in rows),
               "tensor_pipe_active_pct": {"up": rows[0]["tensor_pipe_pct"], "down": rows[1]["tensor_pipe_pct"],
                                          "source": "ncu sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"},
               "launches": rows}, open(dst, "w"), indent=1)


if __name__ == "__main__":
    {"shares": shares, "metrics": metrics, "side": side}[sys.argv[1]](*sys.argv[2:5])
