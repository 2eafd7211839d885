"""ncu report -> markdown summary for profiles/ (headline metrics of every captured kernel + the stall/segment
analysis of tools/ncu_hot.py).  usage: python tools/ncu_md.py report.ncu-rep "title" "command" > profiles/x.md"""
import csv
import subprocess
import sys

rep, title, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
hdr, units = r[0], r[1]
keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_blocks", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum"]
print("# %s\n\ncommand: `%s`\n" % (title, cmd))
for row in r[2:]:
    print("## %s\n" % row[hdr.index("Kernel Name")][:90])
    print("| metric | value | unit |\n|---|---|---|")
    for k in keep:
        if k in hdr:
            print("| %s | %s | %s |" % (k, row[hdr.index(k)], units[hdr.index(k)]))
    print()
hot = subprocess.run([sys.executable, "tools/ncu_hot.py", rep, "2.0"], capture_output=True, text=True).stdout
print("## stall samples and instruction share between barriers (tools/ncu_hot.py)\n\n```")
for line in hot.splitlines():
    if line.startswith("== ") or line.startswith("   gpu__time"):
        continue
    print(line[:200])
print("```")
