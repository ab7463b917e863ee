"""Per-CUDA-source-line stall samples of one kernel of an .ncu-rep (needs -lineinfo + --import-source on)."""
import csv
import io
import subprocess
import sys

rep, kernel = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kernel, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, hdr, res = "?", None, []
for r in rows:
    if len(r) == 2 and r[0] in ("File Path", "File Name"):
        fname = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit():
        i_s, i_ex = hdr.index("# Samples"), hdr.index("Instructions Executed")
        st = sorted(((int(r[i]) if r[i].isdigit() else 0, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h), reverse=True)
        res.append((int(r[i_s]) if r[i_s].isdigit() else 0, int(r[i_ex]) if r[i_ex].isdigit() else 0, fname, int(r[0]), r[1].strip(), st[:2]))
tot_s, tot_i = sum(x[0] for x in res), sum(x[1] for x in res)
print(f"{kernel}: samples {tot_s}, warp-instructions {tot_i}")
for s, ex, f, ln, src, st in sorted(res, reverse=True)[:top]:
    print(f"{s:7d} {100*s/tot_s:5.1f}%  inst {ex:10d} {100*ex/tot_i:5.1f}%  {f}:{ln:<4d} {src[:80]:80s} {st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]}")
