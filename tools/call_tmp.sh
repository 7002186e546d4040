cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x -k "fp32 or gemm" 2>&1 | tail -5
timeout 300 python tools/fp32_check.py 2>&1 | tail -4 | tee gpurun_out/r02b_fp32_check.log
