"""Summarise `ncu --set full` reports (gpurun_out/*.ncu-rep) into a text table for profiles/."""
import csv, io, subprocess, sys

WANT = [("gpu__time_duration.sum", "time_us"), ("dram__bytes_read.sum", "dram_rd_MB"), ("dram__bytes_write.sum", "dram_wr_MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"), ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1_%"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_%"), ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_%"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_%"), ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
        ("launch__occupancy_limit_shared_mem", "occ_lim_smem"), ("launch__waves_per_multiprocessor", "waves")]


def main(paths):
    for path in paths:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        idx = [(lab, hdr.index(m)) for m, lab in WANT if m in hdr]
        ki, gi, bi = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Block Size")
        print(f"## {path}  (ncu --set full --clock-control none; cold cache, replayed)")
        print("kernel | grid | block | " + " | ".join(f"{lab}[{units[i]}]" for lab, i in idx))
        for r in rows[2:]:
            name = r[ki].split("(")[0].replace("void hvx::", "").replace("hvx::", "")
            print(f"{name} | {r[gi]} | {r[bi]} | " + " | ".join(r[i] for _, i in idx))
        print()


if __name__ == "__main__":
    main(sys.argv[1:])
