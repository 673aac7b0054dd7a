#!/bin/bash
# round 2, GPU call 8 (N B200, N = 4 or 8): bench.py on cfg5 with its parity check, then the drop-in demo on all GPUs
cd "$GRAFT_REPO_ROOT" || exit 1
N=${1:-4}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv,noheader | head -8 > gpurun_out/r02_box_n$N.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_n${N}_cfg5.json 2> gpurun_out/r02_bench_n${N}_cfg5.log
echo "bench n$N rc=$?"; python - <<PY
import json
try:
    j=json.load(open('gpurun_out/r02_bench_n${N}_cfg5.json'))
    print('N=$N cfg5 ms', round(j['ms_per_step'],2), 'TF', round(j['value']/1e3,2), 'e2e ms', round(j['e2e']['ms_per_step'],2), '\nparity', j['parity'], '\nsingle', j['single_gpu_same_config'], '\nexchange', j['exchange'], '\nwhole_step', j['roofline']['whole_step'], '\nkernels', {k:(round(v['ms'],1), round(v['tflops'],1)) for k,v in j['roofline']['kernels'].items()}, '\n', j['config']['parallelism'])
except Exception as e: print('ERR', e)
PY
grep -v "ProcessGroupNCCL" gpurun_out/r02_bench_n${N}_cfg5.log | tail -12 | cut -c1-300
if [ "$2" = "demo" ]; then
  python scripts/make_cfg_file.py cfg5 /tmp/cfg5.sfb
  (time SF_DEVSLOT=5380000000 timeout 900 sparse-matrix-factorization-library_b200/Cholesky/Demo/demo /tmp/cfg5.sfb) > gpurun_out/r02_demo_cfg5_${N}gpu.log 2>&1; echo "demo $N GPUs rc=$?"; tail -14 gpurun_out/r02_demo_cfg5_${N}gpu.log
fi
du -sh gpurun_out
