#!/bin/bash
# time the forward groups with every tune/tc_*.so variant of the tensor-core library
for lib in tune/tc_*.so; do
  YF_B200_LIB=$PWD/$lib timeout 200 python tools/profile_groups.py 512x640 256 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); g = d['groups']
print('$lib', 'err %.2e' % d['err'], 'total %.3f' % d['total_ms'], ' '.join('%s %.3f' % (k, g[k]) for k in '${KEYS:-conv5_4 head_5 conv4_1_3 head_4}'.split()))"
done
