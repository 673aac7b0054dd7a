#!/bin/bash
# round 2, GPU call 5 (one B200): tensor-pipe TRSM, ranks sharing one GPU after the implicit-synchronization fixes
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or restatement or compiled_reference or full_size_backward or narrow_blocking or graph_replays") > gpurun_out/r02_pytest_gpu_5a.log 2>&1
echo "pytest a rc=$?"; tail -5 gpurun_out/r02_pytest_gpu_5a.log
(time timeout 900 python -m pytest tests/test_zz_partition_gpu.py -m gpu -q --durations=25) > gpurun_out/r02_pytest_gpu_5b.log 2>&1
echo "pytest b (partition, ranks share the GPU) rc=$?"; tail -32 gpurun_out/r02_pytest_gpu_5b.log
(SFB200_TEST_SHARED_GRAPHS=1 timeout 400 python -m pytest tests/test_zz_partition_gpu.py -m gpu -q -k "sharing_one_gpu") > gpurun_out/r02_pytest_shared_graphs.log 2>&1; echo "shared+graphs rc=$?"; tail -4 gpurun_out/r02_pytest_shared_graphs.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_5.json 2> gpurun_out/r02_bench_5.log; echo "bench rc=$?"
SFB200_TRSM_MMA=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_5_trsm_vector.json 2> gpurun_out/r02_bench_5_trsm_vector.log
python bench.py --config cfg1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_cfg1.json 2> gpurun_out/r02_bench_cfg1.log
python bench.py --config cfg3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_cfg3.json 2> gpurun_out/r02_bench_cfg3.log
python scripts/launch_table.py cfg2 > gpurun_out/r02_launch_table_cfg2_v3.log 2>&1
python scripts/ncu_target.py cfg2 1 > gpurun_out/r02_cfg2_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_trsm_mma" -s 300 -c 2 -o gpurun_out/r02_trsm_mma_v1 python scripts/ncu_target.py cfg2 1 > gpurun_out/ncu3.log 2>&1
for f in gpurun_out/*.ncu-rep; do [ -f "$f" ] && [ $(stat -c %s "$f") -gt 30000000 ] && rm -f "$f"; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_bench_5*.json'))+['gpurun_out/r02_bench_cfg1.json','gpurun_out/r02_bench_cfg3.json']:
    try:
        j=json.load(open(f)); print(f, j['config']['workload'][:5], 'ms', round(j['ms_per_step'],2), 'TF', round(j['value']/1e3,2), 'e2e ms', round(j['e2e']['ms_per_step'],2), {k:(round(v['ms'],1), round(v['tflops'],1)) for k,v in j['roofline']['kernels'].items()})
    except Exception as e: print(f, 'ERR', e)
PY
grep -E "gemm64 |gemm128 K|potrf:|trsm|total" gpurun_out/r02_launch_table_cfg2_v3.log | tail -12
du -sh gpurun_out
