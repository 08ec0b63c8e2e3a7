mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_backward_gpu.py -m gpu -x -q > gpurun_out/pytest27.log 2>&1; echo "== pytest bwd"; tail -4 gpurun_out/pytest27.log
timeout 600 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/bench27.json 2> gpurun_out/bench27.err; echo "== bench"; cat gpurun_out/bench27.json | cut -c1-300; tail -3 gpurun_out/bench27.err
