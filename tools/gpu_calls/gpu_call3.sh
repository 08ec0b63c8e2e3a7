mkdir -p gpurun_out
timeout 300 python tests/tc_probe.py > gpurun_out/tcp_probe.log 2>&1; echo "== probe"; tail -14 gpurun_out/tcp_probe.log | cut -c1-200
timeout 900 python -m pytest tests/test_backward_gpu.py -q --timeout 300 > gpurun_out/pytest_bwd.log 2>&1; echo "== bwd"; tail -12 gpurun_out/pytest_bwd.log | cut -c1-200
timeout 900 python -m pytest tests/test_parity_gpu.py -q --timeout 300 > gpurun_out/pytest_parity.log 2>&1; echo "== parity"; tail -12 gpurun_out/pytest_parity.log | cut -c1-200
timeout 600 python tools/opbench.py --conv tcp --out gpurun_out/opbench_tcp.json > gpurun_out/opbench_tcp.log 2>&1; echo "== opbench tcp"; cat gpurun_out/opbench_tcp.log | cut -c1-200
timeout 600 python tools/opbench.py --conv tcp --unmasked --layers 'c5@32,c6^64,c7@64,c8^128,c9@128,c10^256,c11@256' --out gpurun_out/opbench_tcp_unmasked.json > gpurun_out/opbench_tcp_unmasked.log 2>&1; echo "== opbench tcp unmasked"; grep modconv gpurun_out/opbench_tcp_unmasked.log | cut -c1-200
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_tcp.log 2>&1; echo "== bench"; tail -1 gpurun_out/bench_tcp.log | cut -c1-2500
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'upfirdn2d_fir4|modconv3x3|torgb' -c 8 -o gpurun_out/prof_r1_tcp python tools/ncu_targets.py --conv tcp > gpurun_out/ncu_tcp.log 2>&1; tail -2 gpurun_out/ncu_tcp.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1_tcp.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/b_ncu.log 2>&1; tail -1 gpurun_out/b_ncu.log | cut -c1-200
