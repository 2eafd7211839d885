"""Summarise an ncu report: per kernel headline metrics + the hottest SASS instructions by stall samples.
usage: python tools/ncu_hot.py gpurun_out/x.ncu-rep [min_pct]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
minp = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
hdr = r[0]
keep = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
for row in r[2:]:
    print("==", row[hdr.index("Kernel Name")][:70])
    print("   " + "  ".join("%s=%s" % (k.split(".")[0].replace("launch__", "").replace("smsp__", "").replace("sm__", ""), row[hdr.index(k)]) for k in keep if k in hdr))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
k, data = None, {}
for x in rows:
    if x and x[0] == "Kernel Name":
        k = x[1][:60] + "#%d" % len(data)
        data[k] = []
        continue
    if k is not None:
        data[k].append(x)
seen = set()
for k, v in data.items():
    base = k
    h, body = v[0], v[1:]
    # instruction / sample share of the code between barriers
    ii_, si_, so_ = h.index("Instructions Executed"), h.index("# Samples"), h.index("Source")
    ti_ = h.index("Thread Instructions Executed")
    toti_ = sum(int(x[ii_]) for x in body) or 1
    tots_ = sum(int(x[si_]) for x in body) or 1
    acc = accs = acct = 0
    start = 0
    segs = []
    for i, x in enumerate(body):
        acc += int(x[ii_]); accs += int(x[si_]); acct += int(x[ti_])
        if "BAR.SYNC" in x[so_] or i == len(body) - 1:
            if acc * 100.0 / toti_ > 0.5 or accs * 100.0 / tots_ > 0.5:
                segs.append("[%d-%d] inst %.1f%% smp %.1f%% lanes %.1f" % (start, i, acc * 100.0 / toti_, accs * 100.0 / tots_, acct / max(acc, 1)))
            acc = accs = acct = 0
            start = i + 1
    print("\n## segments of", base)
    for sg in segs:
        print("     " + sg)
    si, ii, ti = h.index("# Samples"), h.index("Instructions Executed"), h.index("Avg. Threads Executed")
    stalls = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(x[si]) for x in body) or 1
    toti = sum(int(x[ii]) for x in body) or 1
    print("\n##", base, "samples", tot, "inst", toti)
    agg = {h[i]: sum(int(x[i]) for x in body) for i in stalls}
    print("  ", {a: round(100 * b / tot, 1) for a, b in sorted(agg.items(), key=lambda z: -z[1])[:7]})
    for idx, x in enumerate(body):
        if int(x[si]) > tot * minp / 100 or int(x[ii]) > toti * minp * 2 / 100:
            top = sorted([(int(x[i]), h[i][6:]) for i in stalls], reverse=True)[:2]
            print("  %4d %5.1f%%s %5.2f%%i thr=%-5s %-58s %s" % (idx, 100 * int(x[si]) / tot, 100 * int(x[ii]) / toti, x[ti], x[1].strip()[:58], top))
