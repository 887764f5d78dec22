"""Turn ncu output into the small text summaries committed under profiles/.

  python tools/summarize_ncu.py report gpurun_out/prof.ncu-rep  > profiles/rNN_<what>.txt
      key metrics of every profiled launch of an `ncu --set full` capture (read with `ncu -i ... --page raw`)
  python tools/summarize_ncu.py launches gpurun_out/launches.csv > profiles/rNN_launches.txt
      per-kernel totals / shares of an `ncu --metrics gpu__time_duration.sum --csv` launch list
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__cluster_size", "cluster"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (legacy mma)"),
    ("sm__inst_executed_pipe_uniform.sum", "uniform-pipe instructions"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("sm__cycles_elapsed.max", "SM cycles"),
]


def short(name: str) -> str:
    name = name.replace("ealdm::", "")
    return name.split("(")[0].replace("void ", "")


def report(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    tensor_cols = [h for h in hdr if "tensor" in h and ("utc" in h.lower() or "tmem" in h.lower())]
    for r in rows[2:]:
        print(f"== {short(r[idx['Kernel Name']])}")
        for k, label in KEYS:
            if k in idx:
                print(f"   {label:34s} {r[idx[k]]} {units[idx[k]]}")
        for k in tensor_cols[:6]:
            print(f"   {k:34s} {r[idx[k]]} {units[idx[k]]}")


def launches(path):
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    i_name, i_metric, i_val, i_unit = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit"))
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for r in rows[1:]:
        if r[i_metric] != "gpu__time_duration.sum":
            continue
        v = float(r[i_val].replace(",", ""))
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[i_unit].split("/")[0], 1.0)
        tot[short(r[i_name])] += v * scale
        cnt[short(r[i_name])] += 1
    total = sum(tot.values())
    print(f"{len(rows) - 1} profiled launches, {total / 1e3:.3f} ms of device time (serialised, cold caches: compare SHARES)")
    for k in sorted(tot, key=lambda k: -tot[k]):
        print(f"  {tot[k] / 1e3:9.3f} ms  {100 * tot[k] / total:5.1f} %  x{cnt[k]:<5d} {k}")


def traffic(path):
    """DRAM / L2 bytes per launch of the tcgen05 conv kernel from a `--metrics dram__bytes_read.sum,
    dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum --csv` launch list -> JSON (bench.py reads
    profiles/r01_conv_tc_traffic.json for roofline.traffic)."""
    import json
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    i_id, i_name, i_metric, i_val, i_unit = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Value",
                                                                   "Metric Unit"))
    unit_scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    tot = defaultdict(float)
    ids, pairs = set(), set()
    for r in rows[1:]:
        if "conv_tc_kernel" not in r[i_name]:
            continue
        ids.add(r[i_id])
        if r[i_name].split("(")[0].rstrip(">").endswith(", 1"):
            pairs.add(r[i_id])
        tot[r[i_metric]] += float(r[i_val].replace(",", "")) * unit_scale.get(r[i_unit], 1.0)
    n = len(ids)
    out = {"kernel": "ealdm::tc::conv_tc_kernel", "launches": n, "launches_as_cta_pairs": len(pairs),
           "dram_read_bytes_total": tot["dram__bytes_read.sum"], "dram_write_bytes_total": tot["dram__bytes_write.sum"],
           "dram_bytes_per_launch": (tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"]) / max(n, 1),
           "l2_bytes_per_launch": tot["lts__t_bytes.sum"] / max(n, 1),
           "duration_ms_total_under_ncu": tot["gpu__time_duration.sum"],
           "command": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum "
                      "--clock-control none --profile-from-start off -k regex:conv_tc_kernel --csv python "
                      "tools/profile_forward.py  (one eager UNet forward, stdiff config, UNet batch 128)"}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    {"report": report, "launches": launches, "traffic": traffic}[sys.argv[1]](sys.argv[2])
