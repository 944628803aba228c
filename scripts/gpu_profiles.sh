#!/bin/bash
# Refresh the committed evidence: smoke, both bench arms, ncu launch list of the bench command, --set full captures
# of the K=50 sweep on C2 (L2-resident n_wk) and of a K=100 sweep on a C4 shard (n_wk streams from HBM).
mkdir -p gpurun_out
python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
python bench.py > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-adlda --no-impute"
$CMD > gpurun_out/plain_l.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv $CMD > gpurun_out/ncu_l.log 2>&1
echo "ncu launches rc=$?"
CMD2="python scripts/profile_sweep.py 50 8"
$CMD2 > gpurun_out/plain50.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 5 -c 2 -o gpurun_out/prof_k50_final -f $CMD2 > gpurun_out/ncu_f.log 2>&1
echo "ncu full k50 rc=$?"
CMD3="python scripts/profile_sweep_c4.py 100 6"
$CMD3 > gpurun_out/plain_c4.log 2>&1 && \
ncu --set full --clock-control none -k regex:sweep_kernel -s 4 -c 2 -o gpurun_out/prof_c4_k100 -f $CMD3 > gpurun_out/ncu_c4.log 2>&1
echo "ncu full c4 rc=$?"; tail -2 gpurun_out/plain_c4.log
