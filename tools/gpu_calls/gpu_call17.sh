mkdir -p gpurun_out
( timeout 60 tools/ubench/umma_bench 1 5 ; echo "--- tmem A"; timeout 60 tools/ubench/umma_bench 1 2 ; echo "--- M=64"; timeout 60 tools/ubench/umma_bench 2 1 ) > gpurun_out/umma_bench.log 2>&1
cat gpurun_out/umma_bench.log
timeout 600 python tests/tc_probe.py > gpurun_out/probe17.log 2>&1; echo "== probe rc=$?"; grep -c "elements off 0/" gpurun_out/probe17.log; grep -v "elements off 0/" gpurun_out/probe17.log | head -20 | cut -c1-250
timeout 900 python tools/opbench.py --conv tcr --prof --out gpurun_out/opbench17.json > gpurun_out/opbench17.log 2>&1; echo "== opbench rc=$?"; grep "modconv\|conv_total" gpurun_out/opbench17.log | cut -c1-200 | tail -30
