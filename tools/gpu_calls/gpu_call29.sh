mkdir -p gpurun_out
E4S_B200_LIB=$PWD/e4s_b200/libe4s_b200_prof.so timeout 600 python tools/opbench.py --conv tcr --prof --layers 'c15@1024,c14^1024,c13@512,c12^512,c6^64' --out gpurun_out/opbench29p.json > gpurun_out/opbench29p.log 2>&1; echo "== prof rc=$?"; grep "prof\|modconv" gpurun_out/opbench29p.log | cut -c1-170
