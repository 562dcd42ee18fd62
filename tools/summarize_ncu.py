"""Turns gpurun_out/*.csv / *.ncu-rep into the small text summaries committed under profiles/."""
import csv
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct", "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct", "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct", "sm__inst_executed_pipe_fp64.avg.pct",
        "sm__inst_executed_pipe_fma.avg.pct", "sm__inst_executed_pipe_alu.avg.pct", "sm__inst_executed_pipe_xu.avg.pct", "sm__inst_executed_pipe_lsu.avg.pct",
        "sm__warps_active.avg.pct", "smsp__issue_active.avg.pct", "smsp__average_warps_issue_stalled", "smsp__inst_executed.sum ",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct", "gpu__dram_throughput.avg.pct", "sm__cycles_active.avg "]


def launches(path, out):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = {}
    for row in csv.DictReader(lines):
        name = row["Kernel Name"]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(row["Metric Value"].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none: python bench.py --steps 3 --warmup 3 (cold-cache, serialised; compare SHARES)\n")
        f.write(f"# total {tot / 1e6:.3f} ms over {sum(a[0] for a in agg.values())} launches\n")
        for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{a[1] / tot:7.4f}  n={a[0]:4d}  total_ns={a[1]:14.0f}  avg_us={a[1] / a[0] / 1e3:10.2f}  {name[:110]}\n")


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on ({rep}); one block per captured launch\n")
        for row in rows[2:]:
            f.write(f"== {row[hdr.index('Kernel Name')]}\n")
            for i, h in enumerate(hdr):
                if any(h.startswith(k.strip()) for k in KEEP) and "per_second" not in h and ".max" not in h and ".min" not in h and ".sum.pct" not in h:
                    if row[i] not in ("0", ""):
                        f.write(f"  {h} = {row[i]} {units[i]}\n")


if __name__ == "__main__":
    # usage: summarize_ncu.py launches <csv> <out.txt> | full <rep> <out.txt>
    if len(sys.argv) == 4 and sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    elif len(sys.argv) == 4 and sys.argv[1] == "full":
        full(sys.argv[2], sys.argv[3])
    else:
        raise SystemExit(__doc__)
