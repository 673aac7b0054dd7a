#!/bin/bash
# round 2, GPU call 16 (one B200, what is left of the budget): the complex path at a size whose top fronts are wide
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 70 python scripts/zbench.py lap3d 64 cholesky ) > gpurun_out/r02_zbench_lap3d64.log 2>&1
echo "zbench lap3d 64 rc=$?"; grep "^{" gpurun_out/r02_zbench_lap3d64.log
( time timeout 50 python scripts/zbench.py cd3d 44 lu ) > gpurun_out/r02_zbench_cd3d44.log 2>&1
echo "zbench cd3d 44 rc=$?"; grep "^{" gpurun_out/r02_zbench_cd3d44.log
python - <<'PY' > gpurun_out/r02_zdemo_prepare.log 2>&1
import sys
sys.path.insert(0, "sparse-matrix-factorization-library_b200")
import sparseframe_b200 as sf
n, r, c, z, sym = sf.matrices.generate_complex("lap3d", 40, hermitian=True)
sf.matrices.write_matrix_market("/tmp/z_lap3d_40.mtx", n, r, c, z, sym)
PY
( time timeout 30 sparse-matrix-factorization-library_b200/Cholesky/Demo/demo /tmp/z_lap3d_40.mtx ) > gpurun_out/r02_dropin_demo_complex_lap3d40.log 2>&1
echo "demo rc=$?"; grep -i "complex\|time\|residual" gpurun_out/r02_dropin_demo_complex_lap3d40.log
