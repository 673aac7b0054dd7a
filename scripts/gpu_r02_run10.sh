#!/bin/bash
# round 2, GPU call 10 (4 B200): partition parameters on cfg5
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 scripts/mgpu_sweep.py cfg5 \
  "SFB200_BALANCE_TOL=1.05" "SFB200_BALANCE_TOL=1.02" "SFB200_BALANCE_TOL=1.01" "SFB200_BALANCE_TOL=1.10" \
  "SFB200_BALANCE_TOL=1.05 SFB200_DIST_PANEL_MIN=1024" "SFB200_BALANCE_TOL=1.05 SFB200_DIST_MIN_M=256" "SFB200_BALANCE_TOL=1.05 SFB200_PUSH=sm" \
  "SFB200_BALANCE_TOL=1.02 SFB200_WEIGHTED_BLOCKS=1" > gpurun_out/r02_mgpu_sweep_n4_cfg5.log 2>&1
echo rc=$?; grep SWEEP gpurun_out/r02_mgpu_sweep_n4_cfg5.log; tail -3 gpurun_out/r02_mgpu_sweep_n4_cfg5.log | cut -c1-300
