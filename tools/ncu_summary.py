"""Key metrics per profiled kernel from an .ncu-rep (ncu --page raw --csv) + per-line hot spots.
usage: python tools/ncu_summary.py gpurun_out/prof_X.ncu-rep [out_pixels_per_launch]"""
import csv, io, subprocess, sys, os
rep = sys.argv[1]
px = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_xu.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    print("====", r[idx["Kernel Name"]][:70], "grid", r[idx.get("Grid Size", 0)], "block", r[idx.get("Block Size", 0)])
    for w in want:
        if w in idx:
            v = r[idx[w]]
            extra = ""
            try:
                if px and units[idx[w]] in ("inst", "", "byte", "Mbyte", "Kbyte", "Gbyte"):
                    f = float(v.replace(",", ""))
                    mult = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9}.get(units[idx[w]], 1.0)
                    extra = "   -> %.3f per px" % (f * mult / px)
            except Exception:
                pass
            print("   %-78s %16s %-8s%s" % (w, v, units[idx[w]], extra))
