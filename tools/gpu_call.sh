#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 90 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "gemm_split" > gpurun_out/pytest_direct.log 2>&1
rc=$?
echo "pytest exit $rc" >> gpurun_out/pytest_direct.log
tail -15 gpurun_out/pytest_direct.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_direct_full.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_direct_full.log
tail -5 gpurun_out/pytest_direct_full.log
for d in 0 1 0 1; do
  BIOM3_RESID_DIRECT=$d timeout 300 python tools/ab_step.py 384 2>/dev/null | tail -1 >> gpurun_out/ab_direct.jsonl
done
cat gpurun_out/ab_direct.jsonl
