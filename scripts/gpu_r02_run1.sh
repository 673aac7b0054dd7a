#!/bin/bash
# round 2, GPU call 1 (one B200): full gpu test tier incl. the new cfg3 / cfg5 / cfg2-entrywise tests, the bench line,
# the CPU-rate probe, and ncu on the HBM-bound side (cfg1: small fronts, assembly, scatter)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r02_box.txt
(time python -m pytest tests -m gpu -x -q --durations=15) > gpurun_out/r02_pytest_gpu_1.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r02_pytest_gpu_1.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_1.json 2> gpurun_out/r02_bench_1.log
echo "bench rc=$?"
python scripts/probes/cpu_ref_probe.py > gpurun_out/r02_cpu_ref_probe.log 2>&1
python scripts/ncu_target.py cfg1 2 > gpurun_out/r02_cfg1_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_cfg1_v1.csv python scripts/ncu_target.py cfg1 1 > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_small|k_assemble|k_gemm" -c 48 -o gpurun_out/r02_cfg1_hbm_v1 python scripts/ncu_target.py cfg1 1 > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/r02_pytest_gpu_1.log; cat gpurun_out/r02_bench_1.json | head -c 1500; cat gpurun_out/r02_cpu_ref_probe.log
