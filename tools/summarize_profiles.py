"""Turn gpurun_out/ scratch (ncu launch lists, .ncu-rep captures) into the small tracked summaries under profiles/.

  python tools/summarize_profiles.py launches gpurun_out/r1_launches.csv profiles/r1_launches_summary.md
  python tools/summarize_profiles.py rep gpurun_out/x.ncu-rep profiles/x.md        (needs ncu on PATH; no GPU)
"""
import csv, io, re, subprocess, sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__cycles_active.avg",
]


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void ", "", name)
    return name[:90]


def launches(src, dst):
    rows = []
    with open(src) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(lines)))
    for r in rd:
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            if r.get("Metric Unit") == "us": v *= 1e3
            if "spin_kernel" in r["Kernel Name"]:          # torch.cuda._sleep: bench.py's head start for the eager profile pass
                continue
            rows.append((short(r["Kernel Name"]), v))
    tot = sum(v for _, v in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for k, v in rows:
        agg[k][0] += 1; agg[k][1] += v
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary of `{src}`\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised launches: compare SHARES).\n\n")
        f.write(f"{len(rows)} launches, {tot/1e6:.3f} ms total device time\n\n| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {v/1e3:.1f} | {100*v/tot:.1f}% | {v/1e3/n:.2f} |\n")
    print("wrote", dst)


def rep(src, dst, note=""):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary of `{src}`\n\n{note}\n\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
            f.write(f"## `{short(d['Kernel Name'])}`  grid {d.get('Grid Size','')} block {d.get('Block Size','')}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in d: f.write(f"| {k} | {d[k]} | {u[k]} |\n")
            f.write("\n")
    print("wrote", dst)


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "launches": launches(sys.argv[2], sys.argv[3])
    else: rep(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
