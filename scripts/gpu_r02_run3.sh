#!/bin/bash
# round 2, GPU call 3 (one B200): full gpu tier on the reworked engine (peer-aware kernels, left-looking sub-blocks), bench
# lines of all single-GPU configs, A/B of the left-looking order, small ncu capture of the HBM-bound side (cfg1)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q --durations=12) > gpurun_out/r02_pytest_gpu_3.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r02_pytest_gpu_3.log
tail -20 gpurun_out/r02_pytest_gpu_3.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_3.json 2> gpurun_out/r02_bench_3.log; echo "bench rc=$?"
SFB200_LEFT_LOOKING=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_3_rightlooking.json 2> gpurun_out/r02_bench_3_rightlooking.log
for cfg in cfg1 cfg4 cfg3 cfg5; do
  python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_$cfg.json 2> gpurun_out/r02_bench_$cfg.log; echo "bench $cfg rc=$?"
done
python scripts/launch_table.py cfg2 > gpurun_out/r02_launch_table_cfg2_v1.log 2>&1
python scripts/ncu_target.py cfg1 2 > gpurun_out/r02_cfg1_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_cfg1_v1.csv python scripts/ncu_target.py cfg1 1 > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none -k regex:"k_small|k_assemble" -c 10 -o gpurun_out/r02_cfg1_small_v1 python scripts/ncu_target.py cfg1 1 > gpurun_out/ncu2.log 2>&1
for f in gpurun_out/*.ncu-rep; do [ -f "$f" ] && [ $(stat -c %s "$f") -gt 30000000 ] && rm -f "$f"; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_bench_*.json')):
    try:
        j=json.load(open(f)); print(f, j['config']['workload'][:5], 'ms', round(j['ms_per_step'],2), 'TF', round(j['value']/1e3,2), 'e2e ms', round(j['e2e']['ms_per_step'],2), {k:(round(v['ms'],1), round(v['tflops'],1)) for k,v in j['roofline']['kernels'].items()})
    except Exception as e: print(f, 'ERR', e)
PY
du -sh gpurun_out
