set -x
python tools/prof_apg.py 2048 6 > gpurun_out/prof_apg_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_vjp_kernel -s 6 -c 3 -o gpurun_out/r2_vjp python tools/prof_apg.py 2048 6 > gpurun_out/prof_apg_ncu.log 2>&1
python tools/prof_driver.py 262144 60 > gpurun_out/prof_step_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mjxb_step_kernel -s 186 -c 1 -o gpurun_out/r2_step python tools/prof_driver.py 262144 60 > gpurun_out/prof_step_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/prof_apg_ncu.log gpurun_out/prof_step_ncu.log
