# round 2, call 1: parity-work-item kernel (UP2) — tests, per-layer A/B, bench; the round-1 probe of the N = 144 design
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c01_pytest.log 2>&1; echo "== pytest rc=$?"; tail -5 gpurun_out/r2c01_pytest.log
UPL="c0^8,c2^16,c4^32,c6^64,c8^128,c10^256,c12^512"
E4S_B200_UP2=0 timeout 300 python tools/opbench.py --only-conv --conv tcr --layers $UPL --out gpurun_out/r2c01_opbench_up2_0.json > gpurun_out/r2c01_opbench_up2_0.log 2>&1; echo "== opbench UP2=0 rc=$?"; grep -o '"kernel": "[^"]*", "ms": [0-9.]*' gpurun_out/r2c01_opbench_up2_0.log
E4S_B200_UP2=1 timeout 300 python tools/opbench.py --only-conv --conv tcr --layers $UPL --out gpurun_out/r2c01_opbench_up2_1.json > gpurun_out/r2c01_opbench_up2_1.log 2>&1; echo "== opbench UP2=1 rc=$?"; grep -o '"kernel": "[^"]*", "ms": [0-9.]*' gpurun_out/r2c01_opbench_up2_1.log
timeout 300 python tools/opbench.py --only-conv --conv tcr --out gpurun_out/r2c01_opbench_auto.json > gpurun_out/r2c01_opbench_auto.log 2>&1; echo "== opbench auto rc=$?"; grep -o '"kernel": "[^"]*", "ms": [0-9.]*' gpurun_out/r2c01_opbench_auto.log; tail -1 gpurun_out/r2c01_opbench_auto.log
timeout 600 python bench.py --no-cpu-baseline --inversion-batch 0 --faceswap-pairs 0 --gpen-batch 0 > gpurun_out/r2c01_bench.json 2> gpurun_out/r2c01_bench.err; echo "== bench rc=$?"; cut -c1-300 gpurun_out/r2c01_bench.json; tail -2 gpurun_out/r2c01_bench.err
(cd tools/ubench && timeout 120 ./upconv_probe > ../../gpurun_out/r2c01_upconv_probe.log 2>&1; echo "== probe rc=$?"; tail -25 ../../gpurun_out/r2c01_upconv_probe.log)
