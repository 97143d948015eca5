set -x
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r2d_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2d_smoke.log 2>&1; tail -2 gpurun_out/r2d_smoke.log
python bench.py > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; tail -c 300 gpurun_out/r2d_bench.err
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r2d_bench_ref.json 2>> gpurun_out/r2d_bench.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ppo --no-sweep --no-apg > gpurun_out/plain_launches.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2d_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ppo --no-sweep --no-apg > gpurun_out/ncu_launches.log 2>&1
python tools/prof_driver.py 262144 60 > gpurun_out/prof_step_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mjxb_step_kernel -s 186 -c 1 -o gpurun_out/r2d_step python tools/prof_driver.py 262144 60 > gpurun_out/prof_step_ncu.log 2>&1
python tools/policy_probe2.py 65536 > gpurun_out/pol_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:policy_act_kernel -s 8 -c 1 -o gpurun_out/r2d_policy python tools/policy_probe2.py 65536 > gpurun_out/pol_ncu.log 2>&1
cat gpurun_out/r2d_pytest.log; ls -la gpurun_out/*.csv gpurun_out/r2d_*.ncu-rep
