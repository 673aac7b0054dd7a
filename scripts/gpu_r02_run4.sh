#!/bin/bash
# round 2, GPU call 4 (one B200): full gpu tier after the fixes (explicit red.global adds, no allocation / host wait between
# rank launches), bench lines, launch table, shared-GPU ranks with CUDA graphs as an extra
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q --durations=12) > gpurun_out/r02_pytest_gpu_4.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r02_pytest_gpu_4.log
tail -22 gpurun_out/r02_pytest_gpu_4.log
(SFB200_TEST_SHARED_GRAPHS=1 timeout 600 python -m pytest tests/test_zz_partition_gpu.py -m gpu -q -k "sharing_one_gpu" ) > gpurun_out/r02_pytest_shared_graphs.log 2>&1; echo "shared+graphs rc=$?"; tail -4 gpurun_out/r02_pytest_shared_graphs.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_4.json 2> gpurun_out/r02_bench_4.log; echo "bench rc=$?"
SFB200_LEFT_LOOKING=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_4_rightlooking.json 2> gpurun_out/r02_bench_4_rightlooking.log
python bench.py --config cfg1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_cfg1.json 2> gpurun_out/r02_bench_cfg1.log
python scripts/launch_table.py cfg2 > gpurun_out/r02_launch_table_cfg2_v2.log 2>&1
SFB200_LEFT_LOOKING=0 python scripts/launch_table.py cfg2 > gpurun_out/r02_launch_table_cfg2_v2_rightlooking.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_bench_*4*.json'))+['gpurun_out/r02_bench_cfg1.json']:
    try:
        j=json.load(open(f)); print(f, j['config']['workload'][:5], 'ms', round(j['ms_per_step'],2), 'TF', round(j['value']/1e3,2), 'e2e ms', round(j['e2e']['ms_per_step'],2), {k:(round(v['ms'],1), round(v['tflops'],1)) for k,v in j['roofline']['kernels'].items()}, j.get('cpu_baseline'))
    except Exception as e: print(f, 'ERR', e)
PY
du -sh gpurun_out
