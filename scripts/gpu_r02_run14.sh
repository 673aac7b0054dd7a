#!/bin/bash
# round 2, GPU call 14 (one B200): ncu launch list of one cfg2 factorization and a full capture of the dominant kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python scripts/ncu_target.py cfg2 1 > gpurun_out/r02_cfg2_plain.log 2>&1 &&
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/r02_launches_cfg2_v1.csv python scripts/ncu_target.py cfg2 1 > gpurun_out/ncu_a.log 2>&1
echo "launch list rc=$?"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_ws" -s 52 -c 1 -o gpurun_out/r02_k_gemm_ws_cfg2 python scripts/ncu_target.py cfg2 1 > gpurun_out/ncu_b.log 2>&1
echo "full capture rc=$?"; tail -3 gpurun_out/ncu_b.log
for f in gpurun_out/*.ncu-rep; do [ -f "$f" ] && [ $(stat -c %s "$f") -gt 40000000 ] && rm -f "$f"; done
wc -l gpurun_out/r02_launches_cfg2_v1.csv; du -sh gpurun_out
