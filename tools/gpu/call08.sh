# round 2, call 8: fast single-thread MMA issue path for region-pure tiles; determinism loop only in deterministic mode
mkdir -p gpurun_out
timeout 300 python tools/opbench.py --only-conv --conv tcr --out gpurun_out/r2c08_opbench_tcr.json > gpurun_out/r2c08_opbench_tcr.log 2>&1; echo "== opbench tcr rc=$?"; grep -o '"kernel": "[^"]*", "ms": [0-9.]*' gpurun_out/r2c08_opbench_tcr.log; tail -1 gpurun_out/r2c08_opbench_tcr.log
timeout 300 python tools/opbench.py --only-conv --conv tch --layers "c10^256,c12^512,c14^1024" --out gpurun_out/r2c08_opbench_tch.json > gpurun_out/r2c08_opbench_tch.log 2>&1; echo "== opbench tch rc=$?"; grep -o '"kernel": "[^"]*", "ms": [0-9.]*' gpurun_out/r2c08_opbench_tch.log
E4S_B200_LIB=$PWD/e4s_b200/libe4s_b200_prof.so timeout 300 python tools/opbench.py --only-conv --conv tcr --prof --layers "c11@256,c13@512,c15@1024" --out gpurun_out/r2c08_prof_tcr.json > gpurun_out/r2c08_prof_tcr.log 2>&1; echo "== prof tcr rc=$?"; grep -v '^{' gpurun_out/r2c08_prof_tcr.log | cut -c1-200
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c08_pytest.log 2>&1; echo "== pytest rc=$?"; tail -6 gpurun_out/r2c08_pytest.log
timeout 300 python tools/enc_bench.py --out gpurun_out/r2c08_enc_bench.json > gpurun_out/r2c08_enc_bench.log 2>&1; echo "== enc bench rc=$?"; tail -1 gpurun_out/r2c08_enc_bench.log | cut -c1-300
timeout 900 python bench.py --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2c08_bench.json 2> gpurun_out/r2c08_bench.err; echo "== bench rc=$?"; cut -c1-300 gpurun_out/r2c08_bench.json; tail -3 gpurun_out/r2c08_bench.err
