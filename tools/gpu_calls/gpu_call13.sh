mkdir -p gpurun_out
timeout 600 python tests/tc_probe.py > gpurun_out/probe13.log 2>&1; echo "== probe"; grep -c "elements off 0/" gpurun_out/probe13.log; grep -v "elements off 0/" gpurun_out/probe13.log | head -20 | cut -c1-250
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest13.log 2>&1; echo "== pytest"; tail -8 gpurun_out/pytest13.log
timeout 600 python bench.py > gpurun_out/bench13.json 2> gpurun_out/bench13.err; echo "== bench"; cat gpurun_out/bench13.json | cut -c1-3000; tail -3 gpurun_out/bench13.err
