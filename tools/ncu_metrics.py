"""`ncu --page raw --csv` dump of ONE forward (30 launches in plan order) -> profiles/r02_ncu_metrics.json: per fused group the DRAM
bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum), tensor-pipe / FMA-pipe / issue utilisation, registers, achieved
warps, the top stall reasons.  bench.py attaches these to its `kernels` entries and takes `roofline.traffic` from here.
    python tools/ncu_metrics.py <raw.csv> <workload tag, e.g. "640x512 b256"> <out.json> [source note]"""
import csv
import json
import sys

GROUPS = ["conv1_4", "res1_1", "conv2_1", "res2_1", "res2_2", "conv3_1", "res3_1", "res3_2", "conv3_4", "res3_3", "res3_4", "res3_5", "res3_6",
          "conv4_1", "res4_1", "res4_2", "res4_3", "res4_4", "conv5_1", "res5_1", "res5_2", "res5_3", "res5_4", "res5_5", "conv5_2", "conv5_4",
          "head_5", "conv4_1_1", "conv4_1_3", "head_4"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}


def num(r, k):
    try:
        return float(r[idx[k]].replace(",", ""))
    except (KeyError, ValueError):
        return None


def byts(r, k):
    v = num(r, k)
    if v is None:
        return None
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[idx[k]], 1)


assert len(data) == len(GROUPS), "expected %d launches (one forward), got %d" % (len(GROUPS), len(data))
stall_keys = [k for k in hdr if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued")]
out = {"workload": sys.argv[2], "source": sys.argv[4] if len(sys.argv) > 4 else "ncu --set full --clock-control none", "kernels": {}}
for g, r in zip(GROUPS, data):
    st = {k.replace("smsp__pcsamp_warps_issue_stalled_", ""): (num(r, k) or 0.0) for k in stall_keys}
    tot = sum(st.values()) or 1.0
    dur = num(r, "gpu__time_duration.sum")
    out["kernels"][g] = {
        "kernel": r[idx["Kernel Name"]].split("(")[0].replace("void yf::", "")[:90],
        "dram_bytes": int((byts(r, "dram__bytes_read.sum") or 0) + (byts(r, "dram__bytes_write.sum") or 0)),
        "duration_us_under_ncu": dur if units[idx["gpu__time_duration.sum"]] in ("usecond", "us") else dur,
        "tensor_pipe_pct": num(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        "fma_pipe_inst_pct": num(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warps_active_pct": num(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "dram_throughput_pct": num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "registers": num(r, "launch__registers_per_thread"),
        "smem_bank_conflict_wavefronts": num(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
        "smem_wavefronts": num(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
        "top_stalls": {k: round(100 * v / tot, 1) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:4]},
    }
json.dump(out, open(sys.argv[3], "w"), indent=1)
print("wrote %s: %d kernels" % (sys.argv[3], len(out["kernels"])))
