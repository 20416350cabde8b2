"""Per-launch summary of an ncu report (the metrics DESIGN.md / the roofline block quote): python tools/ncu_summary.py rep.ncu-rep"""
import csv, subprocess, sys
WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%"),
        ("sm__inst_executed_pipe_uniform.sum", "uniform_inst"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsu_wavefronts_%"),
        ("l1tex__data_bank_conflicts_pipe_lsu.sum", "smem_bank_conflicts"),
        ("lts__t_bytes.sum", "l2_bytes"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dyn_smem"), ("smsp__cycles_active.avg", "cycles")]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
name_i = hdr.index("Kernel Name")
print(f"# ncu --set full --clock-control none; one line per captured launch of {sys.argv[1].split('/')[-1]}")
for r in rows[2:]:
    out = [r[name_i].replace("void ", "")[:70]]
    for key, label in WANT:
        if key in hdr:
            i = hdr.index(key)
            out.append(f"{label}={r[i]}{units[i] if units[i] not in ('', '%') else ''}")
    print("  ".join(out))
