#!/bin/bash
# One 8-GPU gpurun call: bench at N=8 (model sweep + C3 AD-LDA), atlas-scale C4 AD-LDA, north-star target T via
# the public API on all GPUs, the multi-GPU pytest cases.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29521 bench.py --gpus 8 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench n8 rc=$?"
tail -c 1500 gpurun_out/bench_n8.json
timeout 600 $TR --master-port 29522 bench.py --gpus 8 --adlda-only --adlda-shape C4 --steps 5 > gpurun_out/adlda_c4_n8.json 2> gpurun_out/adlda_c4_n8.err; echo "c4 rc=$?"
cat gpurun_out/adlda_c4_n8.json
timeout 600 python scripts/target_sweep.py 2>&1 | grep -v "cisTopic " | tail -3 > gpurun_out/target_T_n8.log; cat gpurun_out/target_T_n8.log
timeout 600 python -m pytest tests -m gpu -x -q -k "adlda or multi" 2>&1 | tail -3
