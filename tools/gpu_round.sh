#!/bin/bash
# One GPU-box session (N = 1): GPU test suite, smoke, bench (reference arm + ours), the 320x256 configuration, and the ncu launch list of
# a short bench run (shares per launch + DRAM bytes). Everything lands in gpurun_out/ (small files only).
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
nvidia-smi > gpurun_out/nvidia-smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/summary.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err
echo "bench ref exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py --res 256x320 --no-cpu-baseline > gpurun_out/bench_256.log 2> gpurun_out/bench_256.err
echo "bench 256 exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --batch 1 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench_b1.log 2> gpurun_out/bench_b1.err
echo "bench b1 exit $?" >> gpurun_out/summary.txt
# every launch of a short bench run with its device time (cold-cache, serialised: compare SHARES)
K='regex:irb_kernel|irbt_kernel|irbtc_kernel|irbtc2_kernel|dwpw_tc_kernel|dense_tc_kernel|dense_ta_kernel|stem_kernel|wstem_kernel|wirb_kernel|upcat_kernel|upcat_tc_kernel|pw_kernel|post_kernel|compact_dets_kernel'
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
  timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -n 3 gpurun_out/pytest.log gpurun_out/smoke.log
