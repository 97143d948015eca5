set -x
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r2_pytest3.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
python bench.py > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; tail -c 300 gpurun_out/r2_bench3.err
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r2_bench3_ref.json 2>> gpurun_out/r2_bench3.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ppo --no-sweep --no-apg > gpurun_out/plain_launches.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ppo --no-sweep --no-apg > gpurun_out/ncu_launches.log 2>&1
python tools/prof_apg.py 2048 6 > gpurun_out/prof_apg_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_vjp_kernel -s 6 -c 1 -o gpurun_out/r2_vjp_final python tools/prof_apg.py 2048 6 > gpurun_out/prof_apg_ncu.log 2>&1
cat gpurun_out/r2_pytest3.log; ls -la gpurun_out/*.csv gpurun_out/r2_vjp_final.ncu-rep
