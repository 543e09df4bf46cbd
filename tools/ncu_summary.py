"""Summarise an `ncu --page raw --csv` dump: one line per kernel launch with the metrics that drive tuning."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}


def g(r, k, d="NA"):
    return r[idx[k]] if k in idx else d


def f(r, k):
    try:
        return float(g(r, k, "0").replace(",", ""))
    except ValueError:
        return 0.0


stall_keys = [k for k in hdr if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued")]
for i, r in enumerate(data):
    name = g(r, "Kernel Name")
    short = name.split("<")[0].replace("void yf::", "").replace("void ", "")
    cfg = name[name.find("<") + 1:][:70]
    st = {k.replace("smsp__pcsamp_warps_issue_stalled_", ""): f(r, k) for k in stall_keys}
    tot = sum(st.values()) or 1
    top = sorted(st.items(), key=lambda kv: -kv[1])[:5]
    dr = f(r, "dram__bytes_read.sum") * (1e3 if units[idx["dram__bytes_read.sum"]] == "Kbyte" else 1e6 if units[idx["dram__bytes_read.sum"]] == "Mbyte" else 1)
    dw = f(r, "dram__bytes_write.sum") * (1e3 if units[idx["dram__bytes_write.sum"]] == "Kbyte" else 1e6 if units[idx["dram__bytes_write.sum"]] == "Mbyte" else 1)
    print("%2d %-12s %-72s dur %8s%s regs %3s blk/SM lim smem %s regs %s | warps_active %5.1f%% issue_active %5.1f%% fma_pipe(inst) %5.1f%% lsu %5.1f%% tensor_pipe %4.1f%% | smem wavefronts lsu %.3g (bank-conflicts %.3g) tensor-core %.3g | dram R/W %.1f/%.1f MB | %s"
          % (i, short, cfg, g(r, "gpu__time_duration.sum"), units[idx["gpu__time_duration.sum"]], g(r, "launch__registers_per_thread"),
             g(r, "launch__occupancy_limit_shared_mem"), g(r, "launch__occupancy_limit_registers"),
             f(r, "sm__warps_active.avg.pct_of_peak_sustained_active"), f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
             f(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active") or f(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
             f(r, "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"), f(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
             f(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"), f(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
             f(r, "l1tex__data_pipe_tc_wavefronts_mem_shared.sum"),
             dr / 1e6, dw / 1e6, " ".join("%s %.0f%%" % (k, 100 * v / tot) for k, v in top)))
