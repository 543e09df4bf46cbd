#!/bin/bash
# weak (256 images per GPU) and fixed-job (BASELINE config 4: 1024 images per step) bench lines on N GPUs: tools/scale_run.sh N
N=$1
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
timeout 600 $R bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
echo "weak exit $?"
timeout 600 $R bench.py --gpus $N --steps 20 --warmup 5 --global-batch 1024 > gpurun_out/bench_n${N}_job1024.log 2> gpurun_out/bench_n${N}_job1024.err
echo "strong exit $?"
python - <<P
import json
for f in ("bench_n$N", "bench_n${N}_job1024"):
    try:
        d = json.loads(open("gpurun_out/%s.log" % f).read().strip().splitlines()[-1])
        print(f, "device %.0f img/s %.3f ms/step | e2e %.0f img/s %.3f ms/step | %s | %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["scaling"], d["gather_check"]))
    except Exception as e:
        print(f, "FAILED", e)
P
tail -3 gpurun_out/bench_n$N.err
