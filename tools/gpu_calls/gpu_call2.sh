mkdir -p gpurun_out
timeout 300 python tests/tc_probe.py > gpurun_out/tc_probe.log 2>&1; echo "== probe"; tail -16 gpurun_out/tc_probe.log
timeout 900 python -m pytest tests/test_backward_gpu.py -q --timeout 300 > gpurun_out/pytest_bwd.log 2>&1; echo "== bwd"; tail -25 gpurun_out/pytest_bwd.log
timeout 600 python -m pytest tests/test_parity_gpu.py -q -k "tc_kernel or tensor_core" --timeout 300 > gpurun_out/pytest_tc.log 2>&1; echo "== tc tests"; tail -8 gpurun_out/pytest_tc.log
timeout 600 python tools/opbench.py --conv simt --out gpurun_out/opbench_simt.json > gpurun_out/opbench_simt.log 2>&1; echo "== opbench simt"; tail -28 gpurun_out/opbench_simt.log
timeout 600 python tools/opbench.py --conv tc --out gpurun_out/opbench_tc.json > gpurun_out/opbench_tc.log 2>&1; echo "== opbench tc"; grep modconv gpurun_out/opbench_tc.log | tail -20
E4S_B200_CONV=auto timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tc.log 2>&1; echo "== bench tc"; tail -2 gpurun_out/bench_tc.log | cut -c1-1500
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'upfirdn2d_fir4|modconv3x3|torgb' -c 8 -o gpurun_out/prof_r1_simt python tools/ncu_targets.py --conv simt > gpurun_out/ncu_simt.log 2>&1; tail -3 gpurun_out/ncu_simt.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'modconv3x3_tc' -c 3 -o gpurun_out/prof_r1_tc python tools/ncu_targets.py --conv tc > gpurun_out/ncu_tc.log 2>&1; tail -3 gpurun_out/ncu_tc.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_simt.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/b_ncu.log 2>&1; tail -2 gpurun_out/b_ncu.log | cut -c1-300
