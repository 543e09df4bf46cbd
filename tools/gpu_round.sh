#!/bin/bash
# One GPU-box session: per-group parity report, GPU test suite, smoke, bench. Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia-smi.txt 2>&1
for cfg in 256x320 512x640 stress; do
  timeout 600 python tools/check_forward.py $cfg 2 > gpurun_out/check_$cfg.log 2>&1
  echo "check $cfg exit $?" >> gpurun_out/summary.txt
done
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/summary.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -5 gpurun_out/check_256x320.log gpurun_out/pytest.log gpurun_out/smoke.log
tail -c 1500 gpurun_out/bench.log
