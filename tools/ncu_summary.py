"""Summarise an .ncu-rep (read here, without a GPU): per-kernel key metrics and top stall sites."""
import csv
import io
import re
import subprocess
import sys
import collections

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "lts__t_sectors_op_red.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")].split("(")[0]}
        for k in KEYS:
            if k in hdr:
                d[k] = r[hdr.index(k)] + " " + units[hdr.index(k)]
        res.append(d)
    return res


def stalls(rep, kernel, top=14):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel, "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        return
    hdr = rows[1]
    i_src, i_s, i_ex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    cols = [(h, hdr.index(h)) for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    data, agg, ops = [], collections.Counter(), collections.Counter()
    for r in rows[2:]:
        try:
            s = int(r[i_s])
        except (ValueError, IndexError):
            continue
        data.append((s, r))
        for h, i in cols:
            if r[i].isdigit():
                agg[h] += int(r[i])
        op = re.sub(r"^@!?U?P\d+\s+", "", r[i_src].strip()).split()[0] if r[i_src].strip() else "?"
        ops[op] += int(r[i_ex]) if r[i_ex].isdigit() else 0
    print(f"  [{kernel}] samples {sum(s for s, _ in data)}; stall reasons: {agg.most_common(6)}")
    print(f"  executed warp-instructions by opcode: {ops.most_common(12)}")
    data.sort(key=lambda x: -x[0])
    for s, r in data[:top]:
        st = sorted([(int(r[i]) if r[i].isdigit() else 0, h) for h, i in cols], reverse=True)[:1]
        print(f"    {s:7d} x{r[i_ex]:>9s}  {r[i_src].strip()[:70]:70s} {st[0][1]}")


if __name__ == "__main__":
    rep = sys.argv[1]
    for d in raw(rep):
        print(d["kernel"])
        for k, v in d.items():
            if k != "kernel":
                print(f"    {k:70s} {v}")
    for k in sys.argv[2:]:
        stalls(rep, k)
