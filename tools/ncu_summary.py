"""Compact per-kernel table from an .ncu-rep (`ncu -i rep --page raw --csv`): time, DRAM bytes, tensor-pipe / L2 / DRAM
utilisation, registers, top warp-stall reasons.   python tools/ncu_summary.py gpurun_out/x.ncu-rep [out.csv]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
want = [
    ("time_us", "gpu__time_duration.sum"),
    ("dram_rd_MB", "dram__bytes_read.sum"),
    ("dram_wr_MB", "dram__bytes_write.sum"),
    ("tensor_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l2_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("issue_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("regs", "launch__registers_per_thread"),
    ("cluster", "launch__cluster_size"),
    ("tma_ld_MB", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum"),
]


def conv(v, u):
    try:
        f = float(v.replace(",", ""))
    except ValueError:
        return v
    u = u.lower()
    if u in ("ns", "nsecond"):
        return f / 1e3
    if u in ("ms", "msecond"):
        return f * 1e3
    if u == "byte":
        return f / 1e6
    if u == "kbyte":
        return f / 1e3
    if u == "gbyte":
        return f * 1e3
    return f


out = []
stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
if not stall_cols:
    stall_cols = [h for h in hdr if "warp_issue_stalled" in h and h.endswith("per_warp_active.pct")]
for d in data:
    rec = {"kernel": d[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("mv::", ""), "grid": d[col["Grid Size"]],
           "block": d[col["Block Size"]]}
    for name, h in want:
        if h in col:
            v = conv(d[col[h]], units[col[h]])
            rec[name] = round(v, 2) if isinstance(v, float) else v
    st = []
    for h in stall_cols:
        try:
            st.append((float(d[col[h]]), h.split("stalled_")[1].split("_per_")[0]))
        except ValueError:
            pass
    st.sort(reverse=True)
    rec["top_stalls"] = " ".join("%s=%.1f" % (n, v) for v, n in st[:3])
    out.append(rec)
keys = list(out[0].keys())
w = csv.DictWriter(open(sys.argv[2], "w", newline="") if len(sys.argv) > 2 else sys.stdout, fieldnames=keys)
w.writeheader()
for r in out:
    w.writerow(r)
