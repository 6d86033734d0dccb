#!/bin/bash
# round-2 closing evidence (one GPU): smoke, parity suite, default bench line, reference arm, ncu launch lists (inference + training),
# DRAM traffic of every gemm_tc launch of one forward.  Outputs under gpurun_out/; summaries go to profiles/ (tools/summarise_ncu.py).
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 1800 python -m pytest tests -m gpu -x -q --durations=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null; echo "bench ref rc=$?"
CENET_B200_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python tools/one_forward.py synapse 64 3 > gpurun_out/ncu_launch.log 2>&1; echo "ncu infer rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_train.csv python tools/one_train_step.py acdc 24 2 > gpurun_out/ncu_launch_train.log 2>&1; echo "ncu train rc=$?"
CENET_B200_GRAPH=0 timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:gemm_tc_kernel --csv --log-file gpurun_out/gemm_traffic.csv python tools/one_forward.py synapse 64 3 > gpurun_out/ncu_traffic.log 2>&1; echo "ncu traffic rc=$?"
