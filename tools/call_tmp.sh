cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
rm -f gpurun_out/r02b_ab_linear2.jsonl
timeout 600 python -m pytest tests -q -m gpu -x -k "attention or small or decode" 2>&1 | tail -8
for v in 0 1 0 1; do BIOM3_LINEAR2=$v timeout 300 python tools/ab_step.py 256 >> gpurun_out/r02b_ab_linear2.jsonl 2>gpurun_out/ab_err.log; done
cat gpurun_out/r02b_ab_linear2.jsonl; tail -3 gpurun_out/ab_err.log
NCU_STEPS=2 timeout 600 ncu --metrics gpu__time_duration.sum,sm__inst_executed.sum --clock-control none -k regex:lin -s 8 -c 4 --csv --log-file gpurun_out/r02b_ncu_lin.csv python tools/ncu_step.py > gpurun_out/ncu_lin.log 2>&1
grep -E "gpu__time|inst_exec" gpurun_out/r02b_ncu_lin.csv | cut -d, -f5,12-15 | cut -c1-150
