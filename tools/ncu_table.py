"""ncu report -> one markdown table row per kernel (the launch with the largest grid of every kernel name).
usage: python tools/ncu_table.py report.ncu-rep "title" "command" > profiles/x.md"""
import csv
import subprocess
import sys

rep, title, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
# a report, or its raw page already exported on the GPU box (`ncu -i x.ncu-rep --page raw --csv > x_raw.csv`)
raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
hdr, units = r[0], r[1]
col = {k: hdr.index(k) for k in hdr}


def f(row, k, scale=1.0):
    try:
        return float(row[col[k]].replace(",", "")) * scale
    except (KeyError, ValueError):
        return float("nan")


def to_bytes(row, k):
    u = units[col[k]].lower()
    return f(row, k, {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0))


best = {}
for row in r[2:]:
    name = row[col["Kernel Name"]].replace("void ", "").replace("<unnamed>::", "").split("(")[0]
    g = f(row, "launch__grid_size")
    if name not in best or g > best[name][0]:
        best[name] = (g, row)
print("# %s\n\ncommand: `%s`\n" % (title, cmd))
print("One row per kernel: its launch with the largest grid (the whole-workload batch step).  `issue` = smsp__issue_active, "
      "`warps` = sm__warps_active (both % of peak), `lanes` = active threads per executed warp instruction, "
      "`DRAM` = dram__bytes_read + dram__bytes_write of the launch, `of peak` = DRAM bytes / time / 6545.9 GB/s.\n")
print("| kernel | grid x block | regs | ms | issue % | warps % | lanes | warp inst | DRAM MB | of peak | smem bank conflicts |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
tot = 0.0
for name, (g, row) in sorted(best.items(), key=lambda kv: -f(kv[1][1], "gpu__time_duration.sum")):
    ms = f(row, "gpu__time_duration.sum")
    if units[col["gpu__time_duration.sum"]] == "us":
        ms /= 1e3
    elif units[col["gpu__time_duration.sum"]] == "ns":
        ms /= 1e6
    dram = to_bytes(row, "dram__bytes_read.sum") + to_bytes(row, "dram__bytes_write.sum")
    tot += ms
    print("| %s | %d x %d | %d | %.3f | %.1f | %.1f | %.1f | %.3g | %.1f | %.1f %% | %.3g |" % (
        name, g, f(row, "launch__block_size"), f(row, "launch__registers_per_thread"), ms,
        f(row, "smsp__issue_active.avg.pct_of_peak_sustained_active"), f(row, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        f(row, "smsp__thread_inst_executed_per_inst_executed.ratio"), f(row, "smsp__inst_executed.sum"), dram / 1e6,
        100 * dram / (ms * 1e-3) / 6545.9e9, f(row, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")))
print("\nsum of the rows: %.2f ms (under ncu: cold caches, serialised; K6/K7/K8 run once per blob)" % tot)
