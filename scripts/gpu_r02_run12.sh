#!/bin/bash
# round 2, GPU call 12 (8 B200): partition tolerance on cfg5 at 8 ranks
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 scripts/mgpu_sweep.py cfg5 \
  "SFB200_BALANCE_TOL=1.05" "SFB200_BALANCE_TOL=1.10" "SFB200_BALANCE_TOL=1.10 SFB200_WEIGHTED_BLOCKS=1" > gpurun_out/r02_mgpu_sweep_n8_cfg5.log 2>&1
echo rc=$?; grep SWEEP gpurun_out/r02_mgpu_sweep_n8_cfg5.log | sort -u; tail -2 gpurun_out/r02_mgpu_sweep_n8_cfg5.log | cut -c1-300
