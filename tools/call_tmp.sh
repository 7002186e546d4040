set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x -k "gemm" > gpurun_out/r02b_pytest_gemm.log 2>&1; echo "exit $?" >> gpurun_out/r02b_pytest_gemm.log; tail -5 gpurun_out/r02b_pytest_gemm.log
TRACE_ONLY=out-proj timeout 300 python tests/tools/gemm_trace.py > gpurun_out/r02b_gemm_trace_epi6.log 2>&1; tail -50 gpurun_out/r02b_gemm_trace_epi6.log
for v in 0 1 2 0 1; do BIOM3_RESID_TMA=$v timeout 300 python tools/ab_step.py 256 >> gpurun_out/r02b_ab_resid_tma.jsonl 2>gpurun_out/ab_err.log; done
cat gpurun_out/r02b_ab_resid_tma.jsonl
