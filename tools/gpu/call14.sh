# round 2, call 14: stacked hi/lo weights (STK) on the small-N layers, experimental TMA-staged chunks for wide masked layers, tests
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c14_pytest.log 2>&1; echo "== pytest rc=$?"; tail -6 gpurun_out/r2c14_pytest.log
for stk in 0 1; do
  E4S_B200_STK=$stk timeout 300 python tools/opbench.py --only-conv --conv tcr --layers "c13@512,c15@1024" --out gpurun_out/r2c14_opbench_stk$stk.json > gpurun_out/r2c14_opbench_stk$stk.log 2>&1; echo "== STK=$stk rc=$?"; grep -o '"kernel": "[^"]*", "ms": [0-9.]*' gpurun_out/r2c14_opbench_stk$stk.log
done
L="c3@16,c4^32,c5@32,c6^64,c7@64,c8^128,c9@128,c11@256"
for xs in 0 1; do
  E4S_B200_XS=$xs timeout 300 python tools/opbench.py --only-conv --conv tcr --layers $L --out gpurun_out/r2c14_opbench_xs$xs.json > gpurun_out/r2c14_opbench_xs$xs.log 2>&1; echo "== XS=$xs rc=$?"; grep -o '"kernel": "[^"]*", "ms": [0-9.]*' gpurun_out/r2c14_opbench_xs$xs.log
done
E4S_B200_LIB=$PWD/e4s_b200/libe4s_b200_prof.so timeout 300 python tools/opbench.py --only-conv --conv tcr --prof --layers "c13@512,c15@1024" --out gpurun_out/r2c14_prof_tcr.json > gpurun_out/r2c14_prof_tcr.log 2>&1; echo "== prof tcr (STK) rc=$?"; grep -v '^{' gpurun_out/r2c14_prof_tcr.log | cut -c1-200
E4S_B200_XS=1 E4S_B200_LIB=$PWD/e4s_b200/libe4s_b200_prof.so timeout 300 python tools/opbench.py --only-conv --conv tcr --prof --layers "c7@64,c9@128" --out gpurun_out/r2c14_prof_xs.json > gpurun_out/r2c14_prof_xs.log 2>&1; echo "== prof tcr (XS=1) rc=$?"; grep -v '^{' gpurun_out/r2c14_prof_xs.log | cut -c1-200
E4S_BENCH_PROFILE_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2c14_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-e2e --inversion-steps 0 --faceswap-pairs 0 --gpen-batch 0 > gpurun_out/r2c14_launches.log 2>&1; echo "== ncu launches rc=$?"; wc -l gpurun_out/r2c14_launches.csv
