#!/bin/bash
# round 2, GPU call 15 (one B200, the last ~3 minutes of the round's budget): the complex path (row f4) -- its GPU tests, the
# smoke check of the real path after the rebuild, a timing of the embedded factorization next to the compiled reference.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/r02_zbox.txt 2>&1
( time timeout 100 python -m pytest tests/test_zzz_complex_gpu.py -m gpu -q --durations=6 -p no:cacheprovider ) > gpurun_out/r02_pytest_complex_gpu.log 2>&1
echo "complex pytest rc=$?"; tail -4 gpurun_out/r02_pytest_complex_gpu.log
( time timeout 40 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r02_smoke_after_complex.log 2>&1
echo "smoke rc=$?"; tail -3 gpurun_out/r02_smoke_after_complex.log
( time timeout 60 python scripts/zbench.py ) > gpurun_out/r02_zbench.log 2>&1
echo "zbench rc=$?"; grep "^{" gpurun_out/r02_zbench.log
( time timeout 40 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "golden" -p no:cacheprovider ) > gpurun_out/r02_pytest_real_golden_after_complex.log 2>&1
echo "real golden rc=$?"; tail -2 gpurun_out/r02_pytest_real_golden_after_complex.log
