#!/bin/bash
# Weak scaling (32 patches per GPU; the SR1 line carries the SR2 sub-record) and strong scaling (global batch 256) of the training step on
# the GPUs of one box.  usage: tools/scaling_runs.sh <tag> [max_gpus]   ->  gpurun_out/<tag>_{weak,strong}_n<N>.json
tag=${1:-scale}; maxn=${2:-8}
mkdir -p gpurun_out
run() {  # n, per-GPU batch, out file, extra flags
    local n=$1 b=$2 out=$3; shift 3
    if [ "$n" = 1 ]; then
        timeout 200 python bench.py --gpus 1 --batch $b --steps 30 --warmup 3 --no-roofline --no-cpu-baseline "$@" > $out 2> $out.err
    else
        timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --batch $b --steps 30 --warmup 3 --no-roofline --no-cpu-baseline "$@" > $out 2> $out.err
    fi
    echo "n=$n batch/gpu=$b rc=$? $(python tools/bench_summary.py $out 2>/dev/null | head -1)"
}
for n in 1 2 4 8; do [ $n -le $maxn ] && run $n 32 gpurun_out/${tag}_weak_n$n.json; done
for n in 1 2 4; do [ $n -le $maxn ] && run $n $((256 / n)) gpurun_out/${tag}_strong256_n$n.json --no-extras; done
