mkdir -p gpurun_out
timeout 600 python tests/tc_probe.py > gpurun_out/probe12.log 2>&1; echo "== probe"; head -24 gpurun_out/probe12.log | cut -c1-300
