"""Distil an .ncu-rep (one kernel launch, --set full) into a small JSON of the counters DESIGN.md / bench.py cite.

    python tools/ncu_summary.py gpurun_out/x_prof.ncu-rep profiles/rNN_step_kernel_ncu_full.json [profiles/step_kernel_dram_bytes.json]
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    res = {"kernel": d.get("Kernel Name", ("", ""))[0]}
    for k in KEYS:
        if k in d:
            res[k] = list(d[k])
    stalls = {}
    for h, (v, u) in d.items():
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h and v:
            stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(v.replace(",", ""))
    res["warp_stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
    json.dump(res, open(out, "w"), indent=1)
    if len(sys.argv) > 3:
        def b(k):
            v, u = d[k]
            return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        json.dump({"dram_bytes_per_launch": b("dram__bytes_read.sum") + b("dram__bytes_write.sum"), "source": rep.split("/")[-1],
                   "kernel": res["kernel"]}, open(sys.argv[3], "w"), indent=1)
    print(json.dumps(res, indent=1)[:1500])


if __name__ == "__main__":
    main()
