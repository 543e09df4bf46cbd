#!/bin/bash
# ncu --set full (with source) of the launches whose kernel name matches $1 in one batch-256 forward (tools/profile_groups.py).
#   tools/ncu_kernels.sh <kernel regex> <launches to skip> <launches to capture> <output name>
# The report comes back as gpurun_out/<name>.ncu-rep; read it here with `ncu -i ... --page raw|source --csv`.
K=${1:-wirb_kernel}; SKIP=${2:-4}; CNT=${3:-4}; OUT=${4:-prof}; RES=${5:-512x640}; B=${6:-256}
mkdir -p gpurun_out
timeout 300 python tools/profile_groups.py $RES $B > gpurun_out/${OUT}_plain.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K" -s $SKIP -c $CNT -f -o gpurun_out/$OUT \
      python tools/profile_groups.py $RES $B > gpurun_out/${OUT}_ncu.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/${OUT}_plain.log | cut -c1-600
ncu -i gpurun_out/$OUT.ncu-rep --page raw --csv > gpurun_out/${OUT}_raw.csv 2>/dev/null && python tools/ncu_summary.py gpurun_out/${OUT}_raw.csv
