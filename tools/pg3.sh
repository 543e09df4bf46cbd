#!/bin/bash
# quick check of the res3 tensor-core groups (GPU box): forward parity tests, then per-group timing at the headline workload
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_forward_gpu.py -m gpu -x -q --timeout 300 > gpurun_out/pg3_pytest.log 2>&1
echo "pytest exit $?"; tail -n 5 gpurun_out/pg3_pytest.log
timeout 300 python tools/profile_groups.py 512x640 256 > gpurun_out/pg.json && python - <<'P'
import json
d=json.load(open("gpurun_out/pg.json")); g=d["groups"]
print(d["total_ms"], d["err"], {k:g[k] for k in ("res3_1","conv3_4","res3_3","conv4_1","res4_1","conv5_1","res5_1","conv4_1_1")})
P
