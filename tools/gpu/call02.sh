# round 2, call 2: H-form up-sampling kernel (modconv_tch.cu) — parity, per-layer A/B against the polyphase forms, bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "tch" > gpurun_out/r2c02_pytest_tch.log 2>&1; echo "== pytest tch rc=$?"; tail -15 gpurun_out/r2c02_pytest_tch.log
UPL="c0^8,c2^16,c4^32,c6^64,c8^128,c10^256,c12^512,c14^1024"
timeout 300 python tools/opbench.py --only-conv --conv tch --layers $UPL --out gpurun_out/r2c02_opbench_tch.json > gpurun_out/r2c02_opbench_tch.log 2>&1; echo "== opbench tch rc=$?"; grep -o '"kernel": "[^"]*", "ms": [0-9.]*' gpurun_out/r2c02_opbench_tch.log; tail -3 gpurun_out/r2c02_opbench_tch.log | cut -c1-300
timeout 300 python tools/opbench.py --only-conv --conv tch --unmasked --layers $UPL --out gpurun_out/r2c02_opbench_tch_unmasked.json > gpurun_out/r2c02_opbench_tch_unmasked.log 2>&1; echo "== opbench tch unmasked rc=$?"; grep -o '"kernel": "[^"]*", "ms": [0-9.]*' gpurun_out/r2c02_opbench_tch_unmasked.log
E4S_B200_UP2=1 timeout 300 python tools/opbench.py --only-conv --conv tcr --unmasked --layers $UPL --out gpurun_out/r2c02_opbench_up2_unmasked.json > gpurun_out/r2c02_opbench_up2_unmasked.log 2>&1; echo "== opbench UP2 unmasked rc=$?"; grep -o '"kernel": "[^"]*", "ms": [0-9.]*' gpurun_out/r2c02_opbench_up2_unmasked.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c02_pytest.log 2>&1; echo "== pytest rc=$?"; tail -5 gpurun_out/r2c02_pytest.log
timeout 600 python bench.py --no-cpu-baseline --inversion-batch 0 --faceswap-pairs 0 --gpen-batch 0 --inversion-steps 0 > gpurun_out/r2c02_bench.json 2> gpurun_out/r2c02_bench.err; echo "== bench rc=$?"; cut -c1-300 gpurun_out/r2c02_bench.json; tail -2 gpurun_out/r2c02_bench.err
