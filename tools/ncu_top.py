"""Summarise an .ncu-rep: headline metrics + top stalled SASS instructions.  python tools/ncu_top.py rep [n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, v = rows[0], rows[-1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "lts__t_sectors_srcunit_tex_op_write.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
for k, val in zip(h, v):
    if k in want or "pipe_tensor" in k and "pct" in k:
        print(f"  {k} = {val}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; data = rows[2:]
ci = {nm: i for i, nm in enumerate(h)}
stall = [nm for nm in h if nm.startswith("stall_") and "Not Issued" not in nm]
tot = sum(int(r[ci["# Samples"]]) for r in data)
print("total samples", tot)
agg = {}
for r in data:
    for nm in stall:
        agg[nm] = agg.get(nm, 0) + int(r[ci[nm]])
print("  by reason:", sorted([(k[6:], x) for k, x in agg.items() if x], key=lambda x: -x[1])[:8])
for r in sorted(data, key=lambda r: -int(r[ci["# Samples"]]))[:n]:
    st = sorted([(nm[6:], int(r[ci[nm]])) for nm in stall if int(r[ci[nm]]) > 0], key=lambda x: -x[1])[:3]
    print(f"{int(r[ci['# Samples']]):6d} {r[ci['Instructions Executed']]:>8s} {r[ci['Source']].strip()[:72]:72s} {st}")
