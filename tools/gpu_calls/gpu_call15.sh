mkdir -p gpurun_out
timeout 600 python tools/opbench.py --conv tcr --prof --layers 'c15@1024,c14^1024,c13@512,c12^512,c11@256,c10^256,c8^128,c6^64,c7@64' --out gpurun_out/opbench15.json > gpurun_out/opbench15.log 2>&1; echo "== opbench rc=$?"; grep -v "^{" gpurun_out/opbench15.log | tail -80
