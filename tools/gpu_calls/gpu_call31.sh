mkdir -p gpurun_out
timeout 600 python tests/tc_probe.py > gpurun_out/probe31.log 2>&1; echo "== probe rc=$?"; grep -c "elements off 0/" gpurun_out/probe31.log; grep -v "elements off 0/" gpurun_out/probe31.log | head -20 | cut -c1-250
timeout 900 python tools/opbench.py --conv tcr --out gpurun_out/opbench31.json > gpurun_out/opbench31.log 2>&1; echo "== opbench rc=$?"; grep "conv_total" gpurun_out/opbench31.log
