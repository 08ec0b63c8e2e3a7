# round 2, call 26: final rehearsal on the last commit - GPU tests, smoke(), default bench line, reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c26_pytest.log 2>&1; echo "== pytest rc=$?"; tail -3 gpurun_out/r2c26_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2c26_smoke.log 2>&1; echo "== smoke rc=$?"; tail -2 gpurun_out/r2c26_smoke.log
timeout 1500 python bench.py > gpurun_out/r2c26_bench_full.json 2> gpurun_out/r2c26_bench_full.err; echo "== full bench rc=$?"; cut -c1-300 gpurun_out/r2c26_bench_full.json
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r2c26_bench_reference.json 2> gpurun_out/r2c26_bench_reference.err; echo "== reference arm rc=$?"; cut -c1-300 gpurun_out/r2c26_bench_reference.json
