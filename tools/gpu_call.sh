#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 90 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "gemm" > gpurun_out/pytest_rd2.log 2>&1
rc=$?
echo "pytest exit $rc" >> gpurun_out/pytest_rd2.log
tail -12 gpurun_out/pytest_rd2.log
if [ $rc -ne 0 ]; then export BIOM3_RESID_DEPTH=1; echo "RD2 FAILED, continuing with depth 1"; fi
timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "attention" > gpurun_out/pytest_attn.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_attn.log
tail -4 gpurun_out/pytest_attn.log
timeout 120 python tools/attn_check.py 2>&1 | grep -E "linear rel_err|variant 3" | tail -8 > gpurun_out/attn_check.log
cat gpurun_out/attn_check.log
if [ $rc -eq 0 ]; then
for rd in 1 2; do
  BIOM3_RESID_DEPTH=$rd BIOM3_EPI_SKIP=0 timeout 120 python tools/gemm_ksweep.py run 2>&1 | grep "split resid" | sed "s/^/rd=$rd /" >> gpurun_out/ksweep_rd.log
done
cat gpurun_out/ksweep_rd.log
fi
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "decode or forward or dropin or full_config or compaction" > gpurun_out/pytest_rd2b.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_rd2b.log
tail -5 gpurun_out/pytest_rd2b.log
for rd in 1 2 1 2; do
  BIOM3_RESID_DEPTH=$rd timeout 300 python tools/ab_step.py 384 2>/dev/null | tail -1 >> gpurun_out/ab_rd.jsonl
done
cat gpurun_out/ab_rd.jsonl
