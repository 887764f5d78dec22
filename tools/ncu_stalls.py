"""Print the stall / pipe / throughput metrics of every launch in an `ncu --set full` report (the numbers DESIGN.md
quotes): python tools/ncu_stalls.py gpurun_out/x.ncu-rep"""
import csv
import io
import subprocess
import sys

WANT = ("gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct",
        "sm__inst_executed_pipe_", "smsp__average_warps_issue_stalled", "sm__throughput.avg.pct",
        "l1tex__data_bank_conflicts", "l1tex__data_pipe_lsu_wavefronts_mem_shared", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "lts__t_bytes.sum", "sm__pipe_tensor", "utcmma", "utchmma", "tmem",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__warps_active", "launch__registers", "sm__warps_active")


def main():
    raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].replace("ealdm::", "")
        print("==", name.split("(")[0])
        for i, h in enumerate(hdr):
            if any(w in h for w in WANT) and r[i] not in ("", "0", "n/a"):
                if "stalled" in h and "per_issue_active" not in h:
                    continue
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                if "stalled" in h and v < 0.05:
                    continue
                print(f"   {h:105s} {r[i]} {units[i]}")


if __name__ == "__main__":
    main()
