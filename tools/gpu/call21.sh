# round 2, call 21: round-end rehearsal on the committed state - full GPU test suite, smoke(), the default bench line (now with
# roofline.static and the TF32-loss-network figure), the resolution / batch sweep (BASELINE configs[4])
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c21_pytest.log 2>&1; echo "== pytest rc=$?"; tail -3 gpurun_out/r2c21_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2c21_smoke.log 2>&1; echo "== smoke rc=$?"; tail -2 gpurun_out/r2c21_smoke.log
timeout 1500 python bench.py > gpurun_out/r2c21_bench_full.json 2> gpurun_out/r2c21_bench_full.err; echo "== full bench rc=$?"; cut -c1-300 gpurun_out/r2c21_bench_full.json
timeout 900 python tools/sweep.py --out gpurun_out/r2c21_sweep.json > gpurun_out/r2c21_sweep.log 2>&1; echo "== sweep rc=$?"; tail -25 gpurun_out/r2c21_sweep.log
