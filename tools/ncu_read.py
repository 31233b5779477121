"""Summarise an .ncu-rep: headline metrics + per-SASS-instruction stall hot spots. usage: ncu_read.py file.ncu-rep [topN]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_fp64.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sectors_op_read.sum", "lts__t_bytes.sum", "sm__cycles_elapsed.max", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed_pipe_fmaheavy.sum", "smsp__inst_executed_pipe_fmalite.sum", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active"]
for i, name in enumerate(h):
    if name in keys or name.startswith("smsp__average_warps_issue_stalled") and name.endswith("per_issue_active.ratio"):
        vals = [r[i] for r in rows[2:]]
        try:
            if name.startswith("smsp__average_warps") and float(vals[0]) < 0.2: continue
        except ValueError: pass
        print(f"{name:84s} {rows[1][i]:10s} {vals}")
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h = rows[hdr[0]]; end = hdr[1] - 1 if len(hdr) > 1 else len(rows)
data = [r for r in rows[hdr[0] + 1:end] if len(r) >= len(h) - 2]
ci = {n: i for i, n in enumerate(h)}
tot = sum(int(r[ci["# Samples"]]) for r in data)
print("SASS instrs", len(data), "samples", tot)
for s in [n for n in h if n.startswith("stall_") and "Not Issued" not in n]:
    v = sum(int(r[ci[s]]) for r in data)
    if v > tot * 0.02: print(f"  {s:24s} {v / tot:.3f}")
print("--- hot instructions")
for r in sorted(data, key=lambda r: -int(r[ci["# Samples"]]))[:topn]:
    st = {s[6:]: int(r[ci[s]]) for s in h if s.startswith("stall_") and "Not Issued" not in s and int(r[ci[s]]) > 0}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"  {data.index(r):5d} {int(r[ci['# Samples']]) / tot:.3f} {r[1].strip()[:60]:60s} {top}")
print("--- samples by 100-instruction region (idx, share, executed)")
for i in range(0, len(data), 100):
    s = sum(int(r[ci["# Samples"]]) for r in data[i:i + 100]); ex = sum(int(r[ci["Instructions Executed"]]) for r in data[i:i + 100])
    if s: print(f"  {i:5d} {s / tot:.3f} {ex}")
