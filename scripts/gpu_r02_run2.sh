#!/bin/bash
# round 2, GPU call 2 (TWO B200): the peer-memory exchange on real NVLink -- in-process group, separate processes over CUDA
# IPC, bench.py under torchrun with its parity check, and the P2P probe
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/r02_box2.txt; nvidia-smi topo -m >> gpurun_out/r02_box2.txt 2>&1
(time timeout 900 python -m pytest tests -m gpu -x -q --durations=10 -k "golden or partitioned or separate_processes or entry_points") > gpurun_out/r02_pytest_gpu_2gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r02_pytest_gpu_2gpu.log
tail -15 gpurun_out/r02_pytest_gpu_2gpu.log
for mode in dma sm; do
SFB200_PUSH=$mode SFB200_BENCH_MGPU_CFG=cfg2 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_n2_cfg2_$mode.json 2> gpurun_out/r02_bench_n2_cfg2_$mode.log
echo "bench n2 $mode rc=$?"; head -c 1200 gpurun_out/r02_bench_n2_cfg2_$mode.json; tail -5 gpurun_out/r02_bench_n2_cfg2_$mode.log
done
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/p2p_atomics scripts/probes/p2p_atomics.cu && timeout 300 /tmp/p2p_atomics > gpurun_out/r02_p2p_atomics.log 2>&1
cat gpurun_out/r02_p2p_atomics.log
du -sh gpurun_out
