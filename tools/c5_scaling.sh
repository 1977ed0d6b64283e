#!/usr/bin/env bash
# Long-video measurement (BASELINE configs[4]: 1000 frames 288x512 swin-5, 9970 directed edges, 70.6 GB of pixel-slot
# planes with flow).  NOT part of the default bench (the default workload is c2).  Run on the GPU box:
#
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/c5_scaling.sh 8'        # 8 frame windows, ~8.8 GB of planes per GPU
#   gpurun          --timeout 1500 -- 'bash tools/c5_scaling.sh 1'       # the same scene on one GPU (fits in 180 GB)
#
# Every rank draws only the edges living on its frame window, on its GPU, row by row while they are packed
# (synth.make_scene_device: the 70 GB of pair predictions never exist on the host).  No end-to-end leg (there is no host
# copy of the inputs to upload).  Output: one bench JSON line in gpurun_out/c5_n<N>.json; its loss_hist_head is what
# tools/c5_compare.py compares between the 1-GPU and the 8-GPU run.
set -euo pipefail
N="${1:-8}"
STEPS="${2:-40}"
mkdir -p gpurun_out
if [ "$N" -gt 1 ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus "$N" --workload c5 --steps "$STEPS" --warmup 3 --no-e2e --no-cpu-baseline \
    > "gpurun_out/c5_n${N}.json" 2> "gpurun_out/c5_n${N}.err"
else
  python bench.py --gpus 1 --workload c5 --steps "$STEPS" --warmup 3 --no-e2e --no-cpu-baseline \
    > "gpurun_out/c5_n1.json" 2> "gpurun_out/c5_n1.err"
fi
tail -c 600 "gpurun_out/c5_n${N}.json"
