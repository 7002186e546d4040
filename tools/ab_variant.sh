#!/bin/bash
# A/B of compile-time variants of the library (python -m biom3_b200.build --variant NAME DEFINES...), selected with BIOM3_LIB:
# parity tests on each variant, then the step time of base and variants interleaved.   bash tools/ab_variant.sh ew8 rp ew8rp
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
VARIANTS="${@:-ew8 rp ew8rp}"
for v in $VARIANTS; do
  BIOM3_LIB=$GRAFT_REPO_ROOT/biom3_b200/libbiom3_b200.$v.so timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "gemm or forward or decode_vs or full_config" > gpurun_out/pytest_$v.log 2>&1
  echo "$v pytest exit $?" | tee -a gpurun_out/pytest_$v.log
  tail -2 gpurun_out/pytest_$v.log
done
for rep in 1 2; do
  for v in base $VARIANTS; do
    if [ $v = base ]; then unset BIOM3_LIB; else export BIOM3_LIB=$GRAFT_REPO_ROOT/biom3_b200/libbiom3_b200.$v.so; fi
    timeout 300 python tools/ab_step.py 384 2>/dev/null | tail -1 | sed "s/^/$v /" >> gpurun_out/ab_variants.jsonl
  done
done
cat gpurun_out/ab_variants.jsonl
