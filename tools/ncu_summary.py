"""Summarise an .ncu-rep (raw + source pages) into a short text report for profiles/."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "sm__inst_executed_pipe_tensor.sum"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("kernel:", d.get("Kernel Name", "?")[:110])
    for k in KEYS:
        if k in d:
            print(f"  {k:75s} {d[k]:>16s} {units[hdr.index(k)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 3:
    hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
    def f(r, k):
        try: return float(r[ix[k]])
        except Exception: return 0.0
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    tot = sum(f(r, "# Samples") for r in data) or 1
    reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    print("  warp-state samples:", int(tot))
    for k, v in sorted(((k, sum(f(r, k) for r in data)) for k in reasons), key=lambda x: -x[1])[:8]:
        print(f"    {k:26s} {v/tot*100:5.1f}%")
    print("  hottest instructions (samples, executed, sass):")
    for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:int(sys.argv[2]) if len(sys.argv) > 2 else 12]:
        print(f"    {int(f(r,'# Samples')):6d} {int(f(r,'Instructions Executed')):9d}  {r[ix['Source']][:80]}")
