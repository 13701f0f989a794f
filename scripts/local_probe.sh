for s in 50 64 128 256; do python scripts/prof_one.py c5 0 $s 2>&1 | tail -1; done
for s in 50 200; do python scripts/prof_one.py c3 0 $s 2>&1 | tail -1; done
timeout 200 ncu --set full --import-source on --clock-control none -k regex:local_steps_kernel -c 1 -f -o gpurun_out/prof_hmc_c3 python scripts/prof_one.py c3 0 64 > gpurun_out/ncu_c3.log 2>&1; tail -1 gpurun_out/ncu_c3.log
timeout 200 ncu --set full --import-source on --clock-control none -k regex:local_steps_kernel -c 1 -f -o gpurun_out/prof_mala_c5 python scripts/prof_one.py c5 0 50 > gpurun_out/ncu_c5.log 2>&1; tail -1 gpurun_out/ncu_c5.log
