#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_shift_full.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_shift_full.log
tail -5 gpurun_out/pytest_shift_full.log
timeout 300 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none --launch-skip 118 --launch-count 7 --csv --log-file gpurun_out/inst_counts.csv python tools/ncu_step.py > /dev/null 2>&1
grep -v "^==" gpurun_out/inst_counts.csv | cut -d, -f5,13- | head -20
for i in 1 2; do
  timeout 300 python tools/ab_step.py 384 2>/dev/null | tail -1 >> gpurun_out/ab_shift.jsonl
done
cat gpurun_out/ab_shift.jsonl
