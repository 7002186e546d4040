#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
BIOM3_TMA_STORE=2 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "gemm or forward or decode_vs_reference" > gpurun_out/pytest_store2.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_store2.log
tail -4 gpurun_out/pytest_store2.log
for st in 1 2; do
  BIOM3_TMA_STORE=$st BIOM3_EPI_SKIP=0 python tools/gemm_ksweep.py run 2>&1 | grep -v "split resid" | sed "s/^/store=$st /" >> gpurun_out/ksweep_store.log
done
cat gpurun_out/ksweep_store.log
for st in 1 2 1 2; do
  BIOM3_TMA_STORE=$st timeout 300 python tools/ab_step.py 384 2>/dev/null | tail -1 >> gpurun_out/ab_store.jsonl
done
cat gpurun_out/ab_store.jsonl
