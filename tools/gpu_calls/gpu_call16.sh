mkdir -p gpurun_out
timeout 600 python tests/tc_probe.py > gpurun_out/probe16.log 2>&1; echo "== probe rc=$?"; grep -c "elements off 0/" gpurun_out/probe16.log; grep -v "elements off 0/" gpurun_out/probe16.log | head -20 | cut -c1-250
timeout 900 python tools/opbench.py --conv tcr --prof --out gpurun_out/opbench16.json > gpurun_out/opbench16.log 2>&1; echo "== opbench rc=$?"; grep "modconv\|conv_total" gpurun_out/opbench16.log | cut -c1-200 | tail -30
