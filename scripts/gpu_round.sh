#!/bin/bash
# One gpurun call: GPU parity tests, smoke, both bench arms, then the ncu launch list of the
# bench command and one --set full capture of the K=50 sweep kernel.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
python bench.py > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench rc=$?"
cat gpurun_out/bench1.json | head -c 3000
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-adlda"
$CMD > gpurun_out/plain_l.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv $CMD > gpurun_out/ncu_l.log 2>&1
echo "ncu launches rc=$?"
CMD2="python scripts/profile_sweep.py 50 8"
$CMD2 > gpurun_out/plain50.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 5 -c 2 -o gpurun_out/prof_k50 -f $CMD2 > gpurun_out/ncu_f.log 2>&1
echo "ncu full rc=$?"
