#!/bin/bash
# A/B of programmatic dependent launch (GPU box): forward parity, then the bench at batch 256 and batch 1 with and without YF_NO_PDL
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_forward_gpu.py tests/test_detect_gpu.py -m gpu -x -q --timeout 600 > gpurun_out/pdl_pytest.log 2>&1
echo "pytest exit $?"; tail -n 3 gpurun_out/pdl_pytest.log
for v in 0 1; do
  if [ $v = 1 ]; then export YF_PDL_MAX=0; else export YF_PDL_MAX=100000; fi
  timeout 600 python bench.py --no-cpu-baseline > gpurun_out/pdl_b256_$v.log 2> gpurun_out/pdl_b256_$v.err
  timeout 600 python bench.py --batch 1 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/pdl_b1_$v.log 2> gpurun_out/pdl_b1_$v.err
  python - <<P
import json
for f in ("gpurun_out/pdl_b256_$v.log", "gpurun_out/pdl_b1_$v.log"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print("PDL_OFF=$v", f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), "lat", d.get("latency"))
    except Exception as e:
        print("PDL_OFF=$v", f, "failed", e)
P
done
