cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -x -k "gemm" 2>&1 | tail -2
for v in "" early "" early; do
  if [ -n "$v" ]; then export BIOM3_LIB=$PWD/biom3_b200/libbiom3_b200.$v.so; else unset BIOM3_LIB; fi
  echo "variant '$v'" >> gpurun_out/r02b_ab_refill.jsonl
  timeout 300 python tools/ab_step.py 256 >> gpurun_out/r02b_ab_refill.jsonl 2>gpurun_out/ab_err.log
done
cat gpurun_out/r02b_ab_refill.jsonl
