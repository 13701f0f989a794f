# Round-2 single-GPU measurement pass (run under gpurun; outputs land in gpurun_out/, summaries are copied to profiles/)
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; tail -c 400 gpurun_out/r02_bench_1gpu.err; head -c 900 gpurun_out/r02_bench_1gpu.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_1gpu_reference_arm.json 2>/dev/null; head -c 600 gpurun_out/r02_bench_1gpu_reference_arm.json
FLOWMC_BENCH_EXTRAS=0 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launch_list.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r02_bench_under_ncu.log 2>&1
FLOWMC_BENCH_EXTRAS=0 timeout 400 ncu --set full --import-source on --clock-control none -k regex:local_steps_kernel -s 3 -c 1 -f -o gpurun_out/r02_mala_c2 python bench.py --steps 1 --warmup 3 > gpurun_out/r02_ncu_mala.log 2>&1; tail -2 gpurun_out/r02_ncu_mala.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:flow_backward_tc_kernel -s 1 -c 1 -f -o gpurun_out/r02_bwdtc_c4 python scripts/prof_flow.py c4 train 16384 > gpurun_out/r02_ncu_bwd.log 2>&1; tail -2 gpurun_out/r02_ncu_bwd.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:flow_tc_kernel -s 2 -c 1 -f -o gpurun_out/r02_flowtc_c4_logprob python scripts/prof_flow.py c4 log_prob > gpurun_out/r02_ncu_fwd.log 2>&1; tail -2 gpurun_out/r02_ncu_fwd.log
for c in c4 c5; do timeout 200 python scripts/prof_train.py $c 16384 > gpurun_out/r02_prof_train_${c}_final.txt 2>&1; timeout 200 python scripts/prof_train.py $c 2048 > gpurun_out/r02_prof_train_${c}_2048_split.txt 2>&1; done
timeout 600 python scripts/bench_sampler.py > gpurun_out/r02_sampler_c5_1gpu.json 2> gpurun_out/r02_sampler.err; cat gpurun_out/r02_sampler_c5_1gpu.json | head -c 1200
ls -la gpurun_out | tail -15
