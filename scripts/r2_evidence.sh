#!/bin/bash
# End-of-round evidence on ONE GPU: full GPU test suite, smoke, both bench arms, ncu launch lists, --set full captures of the
# K = 50 sweep and of the fused sweep + exchange launch.  Outputs under gpurun_out/ (copied to profiles/ afterwards).
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu_end.txt 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/r2_pytest_gpu_end.txt
python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r2_smoke_end.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/r2_smoke_end.log
timeout 600 python bench.py --impl reference > gpurun_out/r2_bench_ref_end.json 2> gpurun_out/r2_bench_ref_end.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/r2_bench_n1_end.json 2> gpurun_out/r2_bench_n1_end.err; echo "bench rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-adlda --no-impute --no-parity --no-c5 --no-target"
$CMD > gpurun_out/plain_l.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bench.csv $CMD > gpurun_out/ncu_l.log 2>&1
echo "ncu launches rc=$?"
CMD3="python scripts/profile_overlap.py 50 6"
$CMD3 > gpurun_out/plain_ov.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_overlap.csv $CMD3 > gpurun_out/ncu_lo.log 2>&1
echo "ncu overlap launches rc=$?"
CMD2="python scripts/profile_sweep.py 50 8"
$CMD2 > gpurun_out/plain50.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 5 -c 1 -o gpurun_out/r2_prof_k50 -f $CMD2 > gpurun_out/ncu_f.log 2>&1
echo "ncu k50 rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:sweep_kernel -s 8 -c 1 -o gpurun_out/r2_prof_overlap_k50 -f $CMD3 > gpurun_out/ncu_ov.log 2>&1
echo "ncu overlap rc=$?"
ls -la gpurun_out/*.ncu-rep
