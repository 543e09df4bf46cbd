#!/bin/bash
# quick per-group timing of the thin groups at the headline workload (GPU box)
python tools/profile_groups.py 512x640 256 > gpurun_out/pg.json; python - <<'P'
import json
d=json.load(open("gpurun_out/pg.json")); g=d["groups"]
print(d["total_ms"], d["err"], {k:g[k] for k in ("conv1_4","res1_1","res2_1","conv3_1")})
P
