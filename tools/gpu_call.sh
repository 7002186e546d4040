#!/bin/bash
# Round-end verification on one B200 (run through gpurun): GPU parity suite, smoke, the bench line, the ncu launch list of
# the bench command, one `--set full` capture of a decode layer's kernels.  Results land in gpurun_out/.
# (compute-sanitizer is closed on this pool: profiles/r02_sanitizer_closed_on_pool.log.)
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_full.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_full.log
tail -4 gpurun_out/pytest_gpu_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "bench exit $?"
cat gpurun_out/bench_final.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
# launches 0-8 set-up, then 101 per step: skip two steps, capture layer 2 of the third (QKV, linear, local, out, FF1, FF2)
timeout 900 ncu --set full --import-source on --clock-control none -s 225 -c 6 -o gpurun_out/step_full -f python tools/ncu_step.py > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
timeout 600 python bench.py --precision fp32 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err
echo "bench fp32 exit $?"
cat gpurun_out/bench_fp32.json | cut -c1-600
