#!/bin/bash
# quick GPU check of the tensor-core library variant: parity of every tap, then per-group timings
export YF_B200_LIB=${YF_B200_LIB:-$PWD/yolo_fastest_b200/libyf_b200.so}
timeout 120 python tools/check_forward.py ${1:-512x640} 3 > gpurun_out/check.log 2>&1; echo "check rc $?"
grep -E "BAD|worst|rror" gpurun_out/check.log | head
timeout 200 python tools/profile_groups.py ${1:-512x640} 256 > gpurun_out/prof.log 2>&1; echo "prof rc $?"
tail -1 gpurun_out/prof.log
