"""Per-CUDA-source-line instruction and stall-sample shares of one kernel in an ncu report.
usage: python tools/ncu_lines.py rep.ncu-rep <kernel-regex> [launch-skip] [min_pct]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
minp = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
# a report, or its source page already exported on the GPU box (tools/prof_src.sh: `--page source --csv --print-source cuda,sass`)
out = open(rep).read() if rep.endswith(".csv") else \
    subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + rx,
                    "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
import io
rows = list(csv.reader(io.StringIO(out)))
fname, hdr, lines = None, None, []
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": func = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0] not in ("", ):
        try: ln = int(r[0])
        except ValueError: continue
        if len(r) != len(hdr) or not r[hdr.index("# Samples")].isdigit(): continue
        lines.append((fname, ln, r))
si, ii, ti = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
tots = sum(int(r[si]) for _, _, r in lines) or 1
toti = sum(int(r[ii]) for _, _, r in lines) or 1
print(func[:100]); print("total inst", toti, "samples", tots)
for f, ln, r in lines:
    i, s, t = int(r[ii]), int(r[si]), int(r[ti])
    if i * 100.0 / toti >= minp or s * 100.0 / tots >= minp:
        print("%-16s %4d  inst %5.1f%%  smp %5.1f%%  lanes %4.1f  | %s" % (f, ln, i * 100.0 / toti, s * 100.0 / tots, t / max(i, 1), r[1].strip()[:110]))
