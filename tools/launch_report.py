"""Turn the ncu launch list of one bench run (tools/gpu_round.sh: gpu__time_duration + dram bytes per launch) into the committed
summaries: profiles/r02_launch_shares.txt (share of device time per fused group, to be compared with bench.py's CUDA-event shares)
and profiles/r02_traffic.json (DRAM bytes per launch, read by bench.py for roofline.traffic).

    python tools/launch_report.py gpurun_out/launches.csv 640x512 256
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GROUPS = ["conv1_4", "res1_1", "conv2_1", "res2_1", "res2_2", "conv3_1", "res3_1", "res3_2", "conv3_4", "res3_3", "res3_4", "res3_5",
          "res3_6", "conv4_1", "res4_1", "res4_2", "res4_3", "res4_4", "conv5_1", "res5_1", "res5_2", "res5_3", "res5_4", "res5_5",
          "conv5_2", "conv5_4", "head_5", "conv4_1_1", "conv4_1_3", "head_4", "post"]


def main():
    path, workload, batch = sys.argv[1], sys.argv[2], int(sys.argv[3])
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    per = {}
    for r in rows[1:]:
        if len(r) < len(hdr) or not r[ix["ID"]].isdigit():
            continue
        d = per.setdefault(int(r[ix["ID"]]), {"name": r[ix["Kernel Name"]]})
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(unit, 1)
        d[r[ix["Metric Name"]]] = v * scale
    launches = [per[k] for k in sorted(per)]
    n = len(GROUPS)
    last = max(i for i, l in enumerate(launches) if "post_kernel" in l["name"])
    step = launches[last - n + 1:last + 1]     # the last complete step: 30 forward groups + the head/NMS kernel
    assert len(step) == n and "stem_kernel" in step[0]["name"], "no complete step in the launch list"      # wstem_kernel / stem_kernel
    tot = sum(l["gpu__time_duration.sum"] for l in step)
    out = ["share of device time among this library's kernels in the LAST step of the ncu launch list (%s, %s batch %d; cold-cache,"
           % (os.path.basename(path), workload, batch), "serialised launches: compare SHARES with bench.py's CUDA-event shares, not absolutes)",
           "%-10s %-44s %10s %7s %12s %12s" % ("group", "kernel", "us", "share", "dram rd MB", "dram wr MB")]
    traffic = {}
    for g, l in zip(GROUPS, step):
        rd, wr = l.get("dram__bytes_read.sum", 0.0), l.get("dram__bytes_write.sum", 0.0)
        traffic[g] = int(rd + wr)
        short = l["name"].split("(")[0].replace("void yf::", "").replace("yf::", "")[:44]
        out.append("%-10s %-44s %10.1f %6.1f%% %12.1f %12.1f" % (g, short, l["gpu__time_duration.sum"], 100 * l["gpu__time_duration.sum"] / tot, rd / 1e6, wr / 1e6))
    out.append("%-10s %-44s %10.1f" % ("total", "", tot))
    open(os.path.join(ROOT, "profiles", "r02_launch_shares.txt"), "w").write("\n".join(out) + "\n")
    json.dump({"workload": "%s b%d" % (workload, batch), "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none",
               "bytes_per_launch": traffic}, open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w"), indent=1)
    print("\n".join(out))


if __name__ == "__main__":
    main()
