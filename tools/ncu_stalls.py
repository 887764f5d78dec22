"""Concise stall / pipe summary of every launch in an `ncu --set full` report (the numbers DESIGN.md quotes):
python tools/ncu_stalls.py gpurun_out/x.ncu-rep"""
import csv
import io
import subprocess
import sys

EXACT = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "smsp__inst_executed.sum",
         "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
         "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
         "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
         "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
         "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
         "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
         "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
         "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
         "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
         "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
         "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
         "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    for r in rows[2:]:
        print("==", r[idx["Kernel Name"]].replace("ealdm::", "").split("(")[0])
        for k in EXACT:
            if k in idx:
                print(f"   {k:100s} {r[idx[k]]} {units[idx[k]]}")
        st = []
        for h in stall:
            try:
                st.append((float(r[idx[h]]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
        st.sort(reverse=True)
        print("   stalls per issue: " + ", ".join(f"{n} {v:.2f}" for v, n in st[:8]))


if __name__ == "__main__":
    main()
