mkdir -p gpurun_out
timeout 600 python tests/tc_probe.py > gpurun_out/probe25.log 2>&1; echo "== probe rc=$?"; grep -c "elements off 0/" gpurun_out/probe25.log; grep -v "elements off 0/" gpurun_out/probe25.log | head -20 | cut -c1-250
timeout 900 python tools/opbench.py --conv tcr --out gpurun_out/opbench25.json > gpurun_out/opbench25.log 2>&1; echo "== opbench rc=$?"; grep "conv_total" gpurun_out/opbench25.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest25.log 2>&1; echo "== pytest"; tail -4 gpurun_out/pytest25.log
timeout 600 python bench.py > gpurun_out/bench25.json 2> gpurun_out/bench25.err; echo "== bench"; cat gpurun_out/bench25.json | cut -c1-1200; tail -3 gpurun_out/bench25.err
