mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_full.log 2>&1; echo "== full gpu suite"; tail -6 gpurun_out/pytest_full.log | cut -c1-200
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.log 2>&1; echo "== bench default"; tail -1 gpurun_out/bench_default.log | cut -c1-3000
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "== bench reference"; tail -1 gpurun_out/bench_reference.log | cut -c1-600
