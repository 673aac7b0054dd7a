#!/bin/bash
# round 2, GPU call 6 (one B200, short): ranks of one process sharing the GPU after staging the values through page-locked memory
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export SFB200_PEER_TIMEOUT_S=6
(time timeout 420 python -m pytest tests/test_zz_partition_gpu.py -m gpu -q --durations=15 -x) > gpurun_out/r02_pytest_gpu_6.log 2>&1
echo "pytest partition rc=$?"; tail -25 gpurun_out/r02_pytest_gpu_6.log
python bench.py --config cfg1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_cfg1_v3.json 2> gpurun_out/r02_bench_cfg1_v3.log
python - <<'PY'
import json
j=json.load(open('gpurun_out/r02_bench_cfg1_v3.json')); print('cfg1 ms', round(j['ms_per_step'],2), {k:(round(v['ms'],2), round(v['tflops'],1), round(v['gbs_algorithmic'])) for k,v in j['roofline']['kernels'].items()})
PY
