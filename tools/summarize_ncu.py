"""Turns ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.
  python tools/summarize_ncu.py launches <launches.csv> <out.md> [title]
  python tools/summarize_ncu.py full <report.ncu-rep> <out.md> [title]"""
import collections
import csv
import io
import re
import subprocess
import sys


def launches(path, out, title):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg, tot = collections.defaultdict(lambda: [0, 0.0]), 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    with open(out, "w") as f:
        f.write(f"# {title}\n\nncu `--metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare "
                f"SHARES, not absolutes).  {sum(n for n, _ in agg.values())} launches, {tot:.1f} us total.\n\n")
        f.write("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k.strip()}` | {n} | {t:.1f} | {100 * t / tot:.1f}% | {t / n:.2f} |\n")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def full(path, out, title):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [(i, h) for i, h in enumerate(hdr) if h in WANT]
    ki = hdr.index("Kernel Name")
    with open(out, "w") as f:
        f.write(f"# {title}\n\n`ncu --set full --clock-control none --import-source on` (one launch per row; dram bytes = "
                f"`dram__bytes_read.sum` / `dram__bytes_write.sum` per launch).\n\n")
        f.write("| kernel | " + " | ".join(f"{h} [{units[i]}]" for i, h in cols) + " |\n")
        f.write("|---|" + "---:|" * len(cols) + "\n")
        for r in data:
            name = re.sub(r"\(.*", "", r[ki]).replace("void nvqa::", "")
            f.write(f"| `{name}` | " + " | ".join(r[i] for i, _ in cols) + " |\n")


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else src
    (launches if mode == "launches" else full)(src, dst, title)
