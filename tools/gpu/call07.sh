# round 2, call 7: full suite (determinism, 1024 oracle parity, loss nets with the new stand-in), loss diagnostics, layer times, bench
mkdir -p gpurun_out
timeout 600 python tools/loss_diag.py > gpurun_out/r2c07_loss_diag.log 2>&1; echo "== loss diag rc=$?"; head -16 gpurun_out/r2c07_loss_diag.log | cut -c1-330
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c07_pytest.log 2>&1; echo "== pytest rc=$?"; tail -12 gpurun_out/r2c07_pytest.log
timeout 300 python tools/opbench.py --only-conv --conv tcr --out gpurun_out/r2c07_opbench_tcr.json > gpurun_out/r2c07_opbench_tcr.log 2>&1; echo "== opbench tcr rc=$?"; grep -o '"kernel": "[^"]*", "ms": [0-9.]*' gpurun_out/r2c07_opbench_tcr.log; tail -1 gpurun_out/r2c07_opbench_tcr.log
timeout 900 python bench.py --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2c07_bench.json 2> gpurun_out/r2c07_bench.err; echo "== bench rc=$?"; cut -c1-300 gpurun_out/r2c07_bench.json; tail -3 gpurun_out/r2c07_bench.err
