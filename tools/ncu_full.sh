#!/bin/bash
# ncu --set full over the 30 library kernels of ONE forward at batch 256 (the headline workload), after a plain run of the same command.
# The report itself stays on the box (size). Back in gpurun_out/: the raw CSV page, the one-line-per-kernel summary, and profiles-ready r02_ncu_metrics.json.
K='regex:irb_kernel|irbt_kernel|irbtc_kernel|irbtc2_kernel|dwpw_tc_kernel|dense_tc_kernel|dense_ta_kernel|stem_kernel|wstem_kernel|wirb_kernel|upcat_kernel|upcat_tc_kernel|pw_kernel|pwpw_kernel'
CMD="python tools/profile_groups.py 512x640 256"
mkdir -p gpurun_out
timeout 600 $CMD > gpurun_out/full_plain.log 2>&1 && \
  timeout 1500 ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 30 --launch-count 30 -f -o /tmp/r02_full $CMD > gpurun_out/full_ncu.log 2>&1
echo "ncu full exit $?"
ncu -i /tmp/r02_full.ncu-rep --page raw --csv > gpurun_out/r02_full_raw.csv 2> gpurun_out/full_raw.err
python tools/ncu_summary.py gpurun_out/r02_full_raw.csv > gpurun_out/r02_ncu_full_summary.txt 2>&1
python tools/ncu_metrics.py gpurun_out/r02_full_raw.csv "640x512 b256" gpurun_out/r02_ncu_metrics.json "ncu --set full --clock-control none, one forward at batch 256 (tools/ncu_full.sh)"
wc -l gpurun_out/r02_ncu_full_summary.txt
