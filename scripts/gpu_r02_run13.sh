#!/bin/bash
# round 2, GPU call 13 (one B200): the final state as the driver will see it -- smoke, full gpu tier, the bench line, other configs
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke.log
(time python -m pytest tests -m gpu -x -q --durations=8) > gpurun_out/r02_pytest_gpu_final.log 2>&1
echo "pytest rc=$?"; tail -14 gpurun_out/r02_pytest_gpu_final.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.log; echo "bench rc=$?"
for cfg in cfg1 cfg4 cfg3; do
  python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_${cfg}_final.json 2> gpurun_out/r02_bench_${cfg}_final.log; echo "bench $cfg rc=$?"
done
python - <<'PY'
import json,glob
for f in ['gpurun_out/r02_bench_final.json']+sorted(glob.glob('gpurun_out/r02_bench_cfg*_final.json')):
    try:
        j=json.load(open(f)); print(f, j['config']['workload'][:5], 'ms', round(j['ms_per_step'],2), 'TF', round(j['value']/1e3,2), 'e2e ms', round(j['e2e']['ms_per_step'],2), j['e2e']['api'][:40], {k:(round(v['ms'],1), round(v['tflops'],1)) for k,v in j['roofline']['kernels'].items()}, 'roof', round(j['roofline']['frac'],3), 'whole', round(j['roofline']['whole_step']['frac'],3), (j.get('cpu_baseline') or {}).get('value'))
    except Exception as e: print(f, 'ERR', e)
PY
