mkdir -p gpurun_out
timeout 600 python tests/tc_probe.py > gpurun_out/probe23.log 2>&1; echo "== probe rc=$?"; grep -c "elements off 0/" gpurun_out/probe23.log; grep -v "elements off 0/" gpurun_out/probe23.log | head -20 | cut -c1-250
timeout 900 python tools/opbench.py --conv tcr --out gpurun_out/opbench23.json > gpurun_out/opbench23.log 2>&1; echo "== opbench rc=$?"; grep "conv_total" gpurun_out/opbench23.log
E4S_B200_LIB=$PWD/e4s_b200/libe4s_b200_prof.so timeout 600 python tools/opbench.py --conv tcr --prof --layers 'c15@1024,c14^1024,c13@512,c12^512' --out gpurun_out/opbench23p.json > gpurun_out/opbench23p.log 2>&1; echo "== prof rc=$?"; grep "prof\|modconv" gpurun_out/opbench23p.log | cut -c1-160
