"""Turn ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.
  summarize_ncu.py launches <launches.csv> <out.md>          per-kernel share of one generation (gpu__time_duration)
  summarize_ncu.py full <file.ncu-rep> <out.md>               key metrics of every captured launch (ncu --set full)"""
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

if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
