"""Turn ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.
  summarize_ncu.py launches <launches.csv> <out.md>          per-kernel share of one generation (gpu__time_duration)
  summarize_ncu.py full <file.ncu-rep> <out.md>               key metrics of every captured launch (ncu --set full)
  summarize_ncu.py traffic <gemm_traffic.csv> <steps.log> <out.md> <out.json>
                                                              DRAM bytes / tensor-pipe activity of every GEMM + conv launch of one UNet
                                                              evaluation (ncu --metrics ...), joined with the shapes of the step log"""
import csv, subprocess, sys, collections, re, io

def short(name):
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void\s+", "", name)
    return name.replace("b200::", "")

def launches(path, out):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.reader(io.StringIO("".join(lines)))
    hdr = next(rd)
    ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for r in rd:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum": continue
        v = float(r[iv].replace(",", ""))
        u = r[iu]
        us = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
        k = short(r[ik]); agg[k][0] += 1; agg[k][1] += us; tot += us
    with open(out, "w") as f:
        f.write("# ncu launch list: one generation's worth of consecutive launches of `bench.py` (eager mode)\n\n")
        f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, (n, us) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("| `%s` | %d | %.2f | %.1f %% |\n" % (k, n, us / 1e3, 100 * us / tot))
        f.write("| **total** | %d | %.2f | |\n" % (sum(n for n, _ in agg.values()), tot / 1e3))

WANT = [
    ("duration us", "gpu__time_duration.sum"),
    ("grid", "launch__grid_size"), ("regs", "launch__registers_per_thread"), ("dyn smem KB", "launch__shared_mem_per_block_dynamic"),
    ("tensor pipe active %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("sm throughput %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("issue active %", "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
    ("xu pipe %", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed"),
    ("dram read MB", "dram__bytes_read.sum"), ("dram write MB", "dram__bytes_write.sum"),
    ("dram throughput %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L2->SM TB/s", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second"),
    ("L2 hit %", "lts__t_sector_hit_rate.pct"),
    ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
]

def full(path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none: %s\n\n" % path.split("/")[-1])
        f.write("| # | kernel | " + " | ".join(n for n, _ in WANT) + " |\n|---|---|" + "---:|" * len(WANT) + "\n")
        for j, r in enumerate(data):
            cells = []
            for n, m in WANT:
                if m not in idx: cells.append("-"); continue
                v, u = r[idx[m]].replace(",", ""), units[idx[m]]
                try:
                    x = float(v)
                    if m.startswith("dram__bytes"):
                        x = x / 1e6 if u == "byte" else x / 1e3 if u == "Kbyte" else x * 1e3 if u == "Gbyte" else x
                    if m == "gpu__time_duration.sum":
                        x = x / 1e3 if u in ("ns", "nsecond") else x * 1e3 if u in ("ms", "msecond") else x
                    if m.endswith("per_second"):
                        x = x / 1e3 if u.startswith("Gbyte") else x
                    if m == "launch__shared_mem_per_block_dynamic":
                        x = x / 1e3 if u == "byte" else x
                    cells.append("%.1f" % x if x < 1e5 else "%.3g" % x)
                except ValueError:
                    cells.append(v)
            f.write("| %d | `%s` | %s |\n" % (j, short(r[idx["Kernel Name"]])[:60], " | ".join(cells)))

def traffic(path, steps_log, out_md, out_json):
    import json
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = collections.OrderedDict()
    for r in csv.DictReader(io.StringIO("".join(lines))):
        d = rows.setdefault(r["ID"], {"name": r["Kernel Name"], "grid": r["Grid Size"]})
        d[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
    tob = lambda x: x[0] * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[x[1]]
    tus = lambda x: x[0] / 1e3 if x[1].startswith("n") else x[0] * 1e3 if x[1].startswith("m") else x[0]
    steps = []
    for l in open(steps_log):
        m = re.match(r"step (\S+)\s+kind\s+(\d+)\s+([\d.]+) us\s+out\[([\d,]+)\] in0\[([\d,]+)\] M(\d+) N(\d+) K(\d+)", l)
        if m and m.group(2) in ("14", "15"):
            steps.append((m.group(1), int(m.group(6)), int(m.group(7)), int(m.group(8)), int(m.group(2))))
    tot_fl_steps = sum(2.0 * s[1] * s[2] * s[3] for s in steps)
    if len(steps) != len(rows):
        # split-K steps launch the GEMM kernel more than once (and 1x1 prologues ride along): no 1:1 join with the step log.
        # Totals stay exact (every launch of the family is in the capture); the table is grouped by kernel and grid instead.
        agg = collections.OrderedDict(); tot_b = tot_us = 0.0
        for v in rows.values():
            us = tus(v["gpu__time_duration.sum"]); by = tob(v["dram__bytes_read.sum"]) + tob(v["dram__bytes_write.sum"])
            tp = v["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"][0]
            a = agg.setdefault((short(v["name"])[:48], v["grid"]), [0, 0.0, 0.0, 0.0]); a[0] += 1; a[1] += us; a[2] += by; a[3] += tp
            tot_b += by; tot_us += us
        with open(out_md, "w") as f:
            f.write("# ncu: every tcgen05 GEMM / implicit-conv launch of one SD1.5 UNet evaluation (batch 16, eager)\n\n")
            f.write("`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active...` "
                    "(durations are cold-cache and serialised). %d launches for %d GEMM / conv steps of the plan (split-K steps launch twice).\n\n" % (len(rows), len(steps)))
            f.write("%d launches, %.2f ms, %.2f GB of DRAM traffic = %.1f MB per launch; %.1f TFLOP\n\n" % (len(rows), tot_us / 1e3, tot_b / 1e9, tot_b / 1e6 / len(rows), tot_fl_steps / 1e12))
            f.write("| kernel | grid | launches | us each | tensor pipe % | DRAM MB each |\n|---|---:|---:|---:|---:|---:|\n")
            for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
                n = a[0]
                f.write("| `%s` | %s | %d | %.1f | %.1f | %.1f |\n" % (k[0], k[1], n, a[1] / n, a[3] / n, a[2] / n / 1e6))
        json.dump({"what": "dram__bytes_read.sum + dram__bytes_write.sum over all tcgen05 GEMM/conv launches of one SD1.5 batch-16 UNet evaluation (ncu, one capture)",
                   "launches": len(rows), "dram_bytes_total": tot_b, "dram_bytes_per_launch": tot_b / len(rows), "ncu_time_us_total": tot_us, "flop_total": tot_fl_steps},
                  open(out_json, "w"), indent=1)
        return
    agg = collections.OrderedDict(); tot_b = tot_us = tot_fl = 0.0
    for s, v in zip(steps, rows.values()):
        us = tus(v["gpu__time_duration.sum"]); by = tob(v["dram__bytes_read.sum"]) + tob(v["dram__bytes_write.sum"])
        tp = v["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"][0]
        fl = 2.0 * s[1] * s[2] * s[3]
        a = agg.setdefault(s, [0, 0.0, 0.0, 0.0]); a[0] += 1; a[1] += us; a[2] += by; a[3] += tp
        tot_b += by; tot_us += us; tot_fl += fl
    with open(out_md, "w") as f:
        f.write("# ncu: every tcgen05 GEMM / implicit-conv launch of one SD1.5 UNet evaluation (batch 16, eager)\n\n")
        f.write("`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active...` "
                "(durations are cold-cache and serialised).\n\n")
        f.write("%d launches, %.2f ms, %.2f GB of DRAM traffic = %.1f MB per launch; %.1f TFLOP\n\n" % (len(rows), tot_us / 1e3, tot_b / 1e9, tot_b / 1e6 / len(rows), tot_fl / 1e12))
        f.write("| op | M | N | K | launches | us each | TFLOP/s | tensor pipe % | DRAM MB each |\n|---|---:|---:|---:|---:|---:|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            n = a[0]
            f.write("| %s | %d | %d | %d | %d | %.1f | %.0f | %.1f | %.1f |\n" % (k[0], k[1], k[2], k[3], n, a[1] / n, 2.0 * k[1] * k[2] * k[3] / (a[1] / n) / 1e6, a[3] / n, a[2] / n / 1e6))
    json.dump({"what": "dram__bytes_read.sum + dram__bytes_write.sum over all tcgen05 GEMM/conv launches of one SD1.5 batch-16 UNet evaluation (ncu, one capture)",
               "launches": len(rows), "dram_bytes_total": tot_b, "dram_bytes_per_launch": tot_b / len(rows), "ncu_time_us_total": tot_us, "flop_total": tot_fl},
              open(out_json, "w"), indent=1)


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
