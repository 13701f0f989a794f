set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest7.log 2>&1; tail -3 gpurun_out/pytest7.log
timeout 600 python bench.py > gpurun_out/bench4.json 2> gpurun_out/bench4.err; tail -c 600 gpurun_out/bench4.err; head -c 700 gpurun_out/bench4.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench4_ref.json 2>/dev/null
timeout 300 python scripts/bench_flow.py > gpurun_out/bench_flow_stdout.log 2>&1; cat gpurun_out/bench_flow_stdout.log | cut -c1-220
timeout 600 python scripts/bench_sampler.py > gpurun_out/sampler_c5_1gpu_v2.json 2> gpurun_out/sampler_v2.err; cat gpurun_out/sampler_c5_1gpu_v2.json
timeout 300 ncu --set full --import-source on --clock-control none -k regex:flow_tc_kernel -s 2 -c 1 -f -o gpurun_out/prof_flowtc_c4_v4 python scripts/prof_flow.py c4 log_prob > gpurun_out/ncu_flowtc4.log 2>&1; tail -2 gpurun_out/ncu_flowtc4.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:flow_backward_tc_kernel -s 1 -c 1 -f -o gpurun_out/prof_bwdtc_c4_v3 python scripts/prof_flow.py c4 train 16384 > gpurun_out/ncu_bwd3.log 2>&1; tail -2 gpurun_out/ncu_bwd3.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:flow_tc_kernel -s 1 -c 1 -f -o gpurun_out/prof_flowtc_train_c4 python scripts/prof_flow.py c4 train 16384 > gpurun_out/ncu_fwdtrain.log 2>&1; tail -2 gpurun_out/ncu_fwdtrain.log
FLOWMC_BENCH_EXTRAS=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench2.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launch_train3.csv python scripts/prof_flow.py c4 train 16384 > /dev/null 2>&1
ls -la gpurun_out | tail -12
