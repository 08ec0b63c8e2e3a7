mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest26.log 2>&1; echo "== pytest"; tail -4 gpurun_out/pytest26.log
timeout 600 python bench.py > gpurun_out/bench26.json 2> gpurun_out/bench26.err; echo "== bench"; cat gpurun_out/bench26.json | cut -c1-600; tail -3 gpurun_out/bench26.err
