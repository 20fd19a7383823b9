"""ncu_source_lines.py <report.ncu-rep> [sass|cuda] -- per SASS instruction (or CUDA source line): warp stall samples (all / not issued) and
executed instructions, top 45 lines (ncu --page source --print-source cuda --csv).  Run on the GPU box; prints text."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", (sys.argv[2] if len(sys.argv) > 2 else "sass"), "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr_i = next((i for i, r in enumerate(rows) if "Source" in r and any("Samp" in c for c in r)), None)
if hdr_i is None:
    print("no source table; first lines:"); print("\n".join(out.splitlines()[:15])); sys.exit(0)
hdr = rows[hdr_i]
col = lambda name: next((i for i, c in enumerate(hdr) if c.strip() == name), None)
c_src, c_all, c_ni, c_ex = col("Source"), col("Warp Stall Sampling (All Samples)"), col("Warp Stall Sampling (Not-issued Samples)"), col("Instructions Executed")
if c_all is None:
    c_all = next((i for i, c in enumerate(hdr) if "Sampling" in c and "All" in c), None)
    c_ni = next((i for i, c in enumerate(hdr) if "Sampling" in c and "Not" in c), None)
c_ln = col("#") if col("#") is not None else 0
print("columns:", [hdr[i] for i in (c_ln, c_src, c_all, c_ni, c_ex) if i is not None])
data = []
for r in rows[hdr_i + 1:]:
    if len(r) != len(hdr):
        continue
    try:
        a = float(r[c_all] or 0)
    except ValueError:
        continue
    data.append((a, float(r[c_ni] or 0) if c_ni is not None else 0, float(r[c_ex] or 0) if c_ex is not None else 0, r[c_ln], r[c_src].strip()[:130]))
tot = sum(d[0] for d in data) or 1
print("total samples %d" % tot)
for a, ni, ex, ln, src in sorted(data, reverse=True)[:60]:
    print("%6.2f%%  all=%-7d notissued=%-7d exec=%-9d L%-5s %s" % (100 * a / tot, a, ni, ex, ln, src))
