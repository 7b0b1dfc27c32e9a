"""Summarise an ncu report (one kernel) into profiles/<name>.json + a launch-list markdown.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_bf16_render_512.json [launches.csv]

Reads the report here (no GPU needed) with `ncu -i ... --page raw --csv` and keeps the metrics the
roofline is judged on: duration, tensor-pipe activity, DRAM traffic, issue activity, registers,
shared-memory wavefronts/conflicts, stall mix.  With a launch list (the
`--metrics gpu__time_duration.sum` pass) it also writes the per-kernel time shares."""
import csv
import io
import json
import subprocess
import sys
from collections import defaultdict

KEEP = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "lts__t_sector_hit_rate.pct", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "launch__cluster_size", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.per_cycle_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]
STALLS = "smsp__pcsamp_warps_issue_stalled_"


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    kernels = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")].split("(")[0], "metrics": {}, "stall_samples": {}}
        for k, v, u in zip(hdr, r, units):
            if k in KEEP:
                d["metrics"][k] = {"value": v, "unit": u}
            elif k.startswith(STALLS) and not k.endswith("_not_issued"):
                try:
                    d["stall_samples"][k[len(STALLS):]] = int(float(v))
                except ValueError:
                    pass
        tot = sum(d["stall_samples"].values()) or 1
        d["stall_share_pct"] = {k: round(100.0 * v / tot, 1) for k, v in sorted(d["stall_samples"].items(), key=lambda kv: -kv[1])[:8]}
        del d["stall_samples"]
        kernels.append(d)
    res = {"report": rep.split("/")[-1], "command": "ncu --set full --clock-control none --import-source on", "kernels": kernels}
    if len(sys.argv) > 3:
        by = defaultdict(lambda: [0, 0.0])
        for r in csv.DictReader(l for l in open(sys.argv[3]) if l.startswith('"')):
            if r.get("Metric Name") == "gpu__time_duration.sum":
                name = r["Kernel Name"].split("(")[0]
                by[name][0] += 1
                by[name][1] += float(r["Metric Value"].replace(",", "")) / 1e6
        tot = sum(v[1] for v in by.values()) or 1.0
        res["launch_list"] = {"source": sys.argv[3].split("/")[-1],
                              "note": "cold-cache, serialised per-launch times (ncu --metrics gpu__time_duration.sum --clock-control none): compare shares, not absolutes",
                              "kernels": [{"kernel": k, "launches": v[0], "ms": round(v[1], 3), "share_pct": round(100 * v[1] / tot, 2)}
                                          for k, v in sorted(by.items(), key=lambda kv: -kv[1][1])]}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1)[:3000])


if __name__ == "__main__":
    main()
