# round 2, call 4: elect.sync MMA issue (tcr, tch, dgrad) + 8-warp packed-math epilogue of the H-form kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c04_pytest.log 2>&1; echo "== pytest rc=$?"; tail -8 gpurun_out/r2c04_pytest.log
timeout 300 python tools/opbench.py --only-conv --conv tcr --out gpurun_out/r2c04_opbench_tcr.json > gpurun_out/r2c04_opbench_tcr.log 2>&1; echo "== opbench tcr rc=$?"; grep -o '"kernel": "[^"]*", "ms": [0-9.]*' gpurun_out/r2c04_opbench_tcr.log; tail -1 gpurun_out/r2c04_opbench_tcr.log
UPL="c4^32,c6^64,c8^128,c10^256,c12^512,c14^1024"
timeout 300 python tools/opbench.py --only-conv --conv tch --layers $UPL --out gpurun_out/r2c04_opbench_tch.json > gpurun_out/r2c04_opbench_tch.log 2>&1; echo "== opbench tch rc=$?"; grep -o '"kernel": "[^"]*", "ms": [0-9.]*' gpurun_out/r2c04_opbench_tch.log
E4S_B200_LIB=$PWD/e4s_b200/libe4s_b200_prof.so timeout 300 python tools/opbench.py --only-conv --conv tch --prof --layers "c10^256,c12^512,c14^1024" --out gpurun_out/r2c04_prof_tch.json > gpurun_out/r2c04_prof_tch.log 2>&1; echo "== prof tch rc=$?"; grep -v '^{' gpurun_out/r2c04_prof_tch.log | cut -c1-200
E4S_B200_LIB=$PWD/e4s_b200/libe4s_b200_prof.so timeout 300 python tools/opbench.py --only-conv --conv tcr --prof --layers "c7@64,c9@128,c11@256,c13@512,c15@1024" --out gpurun_out/r2c04_prof_tcr.json > gpurun_out/r2c04_prof_tcr.log 2>&1; echo "== prof tcr rc=$?"; grep -v '^{' gpurun_out/r2c04_prof_tcr.log | cut -c1-200
timeout 600 python bench.py --no-cpu-baseline --inversion-batch 0 --faceswap-pairs 0 --gpen-batch 0 > gpurun_out/r2c04_bench.json 2> gpurun_out/r2c04_bench.err; echo "== bench rc=$?"; cut -c1-300 gpurun_out/r2c04_bench.json; tail -2 gpurun_out/r2c04_bench.err
