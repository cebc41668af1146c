#!/bin/bash
# N-GPU session (gpurun --gpus N): frame-sharded bench (the contract line) and the band-sharded (--tiled) bench over real NCCL.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi_gpus.txt 2>&1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
echo "### frame-sharded x$N"; timeout 600 $T bench.py --gpus $N --steps 5 --warmup 3 --batch 8 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "exit $?"; tail -1 gpurun_out/bench_n$N.log | cut -c1-400
echo "### tiled 8K x$N"; timeout 600 $T bench.py --gpus $N --steps 5 --warmup 3 --batch 2 --width 7680 --height 4320 --tiled --no-cpu-baseline > gpurun_out/bench_tiled8k_n$N.log 2> gpurun_out/bench_tiled8k_n$N.err; echo "exit $?"; tail -1 gpurun_out/bench_tiled8k_n$N.log | cut -c1-400
echo "### tiled 8K x1 (same workload, one GPU)"; timeout 600 python bench.py --steps 5 --warmup 3 --batch 2 --width 7680 --height 4320 --tiled --no-cpu-baseline > gpurun_out/bench_tiled8k_n1.log 2> gpurun_out/bench_tiled8k_n1.err; echo "exit $?"; tail -1 gpurun_out/bench_tiled8k_n1.log | cut -c1-400
echo "### tiled parity x$N (band-sharded over NCCL == single GPU)"; timeout 600 $T tools/tiled_parity_check.py > gpurun_out/tiled_parity_n$N.log 2>&1; echo "exit $?"; tail -3 gpurun_out/tiled_parity_n$N.log
tail -n 3 gpurun_out/bench_n$N.err gpurun_out/bench_tiled8k_n$N.err; true
