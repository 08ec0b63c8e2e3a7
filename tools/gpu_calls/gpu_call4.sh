mkdir -p gpurun_out
timeout 300 python tests/tc_probe.py > gpurun_out/tcq_probe.log 2>&1; echo "== probe"; tail -18 gpurun_out/tcq_probe.log | cut -c1-200
timeout 900 python -m pytest tests/test_backward_gpu.py -q --timeout 300 > gpurun_out/pytest_bwd.log 2>&1; echo "== bwd"; tail -8 gpurun_out/pytest_bwd.log | cut -c1-200
timeout 900 python -m pytest tests/test_parity_gpu.py -q --timeout 300 > gpurun_out/pytest_parity.log 2>&1; echo "== parity"; tail -12 gpurun_out/pytest_parity.log | cut -c1-200
timeout 600 python tools/opbench.py --conv tcq --out gpurun_out/opbench_tcq.json > gpurun_out/opbench_tcq.log 2>&1; echo "== opbench tcq"; grep -E "modconv|conv_total" gpurun_out/opbench_tcq.log | cut -c1-200
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_tcq.log 2>&1; echo "== bench"; tail -1 gpurun_out/bench_tcq.log | cut -c1-1200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'modconv3x3' -c 5 -o gpurun_out/prof_r1_tcq python tools/ncu_targets.py --conv tcq > gpurun_out/ncu_tcq.log 2>&1; tail -2 gpurun_out/ncu_tcq.log
