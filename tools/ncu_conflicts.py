"""Shared-memory bank-conflict hot lines of one kernel in an ncu report.
usage: python tools/ncu_conflicts.py rep.ncu-rep <kernel-regex> [launch-skip] [min-excess]"""
import csv, subprocess, io, sys
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
mn = int(sys.argv[4]) if len(sys.argv) > 4 else 20000
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + rx,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
hdr = None
for r in csv.reader(io.StringIO(out)):
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        ex, tot = hdr.index("L1 Wavefronts Shared Excessive"), hdr.index("L1 Wavefronts Shared")
        if r[ex].isdigit() and int(r[ex]) > mn:
            print(r[0], r[1].strip()[:100], "| excess", r[ex], "of", r[tot])
