#!/bin/bash
# ncu --set full over the 31 library kernels of one timed bench step (batch 64), summarised by tools/ncu_summary.py.
# The report stays on the GPU box (size); only the raw CSV page and the summary come back in gpurun_out/.
K='regex:irb_kernel|irbtc_kernel|irbtc2_kernel|dwpw_tc_kernel|dense_kernel|dense_tc_kernel|stem_kernel|upcat_kernel|upcat_tc_kernel|pw_kernel|post_kernel'
CMD="python bench.py --batch 64 --steps 1 --warmup 3 --no-cpu-baseline"
mkdir -p gpurun_out
timeout 600 $CMD > gpurun_out/full_plain.log 2>&1 && \
  timeout 1500 ncu --set full --clock-control none -k "$K" --launch-skip 93 --launch-count 31 -f -o /tmp/r01_full $CMD > gpurun_out/full_ncu.log 2>&1
echo "ncu full exit $?"
ncu -i /tmp/r01_full.ncu-rep --page raw --csv > gpurun_out/full_raw.csv 2> gpurun_out/full_raw.err
python tools/ncu_summary.py gpurun_out/full_raw.csv > gpurun_out/full_summary.txt 2>&1
wc -l gpurun_out/full_summary.txt
