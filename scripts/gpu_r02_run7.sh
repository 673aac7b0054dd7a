#!/bin/bash
# round 2, GPU call 7 (TWO B200): in-process groups with a GPU per rank (tests + the drop-in demo on cfg2), streamed copy-out of
# partitioned plans in bench.py's e2e leg
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_zz_partition_gpu.py -m gpu -q --durations=15) > gpurun_out/r02_pytest_gpu_7_2gpu.log 2>&1
echo "pytest partition (2 GPUs) rc=$?"; tail -18 gpurun_out/r02_pytest_gpu_7_2gpu.log
python scripts/make_cfg_file.py cfg2 /tmp/cfg2.sfb
(time SF_DEVSLOT=5380000000 timeout 600 sparse-matrix-factorization-library_b200/Cholesky/Demo/demo /tmp/cfg2.sfb) > gpurun_out/r02_demo_cfg2_2gpu.log 2>&1; echo "demo 2 GPUs rc=$?"; tail -12 gpurun_out/r02_demo_cfg2_2gpu.log
(time SF_NUM_GPU=1 SF_DEVSLOT=5380000000 timeout 600 sparse-matrix-factorization-library_b200/Cholesky/Demo/demo /tmp/cfg2.sfb) > gpurun_out/r02_demo_cfg2_1gpu.log 2>&1; echo "demo 1 GPU rc=$?"; tail -8 gpurun_out/r02_demo_cfg2_1gpu.log
SFB200_BENCH_MGPU_CFG=cfg2 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_n2_cfg2_v2.json 2> gpurun_out/r02_bench_n2_cfg2_v2.log
echo "bench n2 rc=$?"; python - <<'PY'
import json
try:
    j=json.load(open('gpurun_out/r02_bench_n2_cfg2_v2.json'))
    print('N=2 cfg2 ms', round(j['ms_per_step'],2), 'TF', round(j['value']/1e3,2), 'e2e ms', round(j['e2e']['ms_per_step'],2), j['parity'], j['single_gpu_same_config'], j['exchange'])
except Exception as e: print('ERR', e)
PY
tail -4 gpurun_out/r02_bench_n2_cfg2_v2.log | cut -c1-400
du -sh gpurun_out
