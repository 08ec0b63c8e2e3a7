mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest21.log 2>&1; echo "== pytest"; tail -6 gpurun_out/pytest21.log
timeout 900 python tools/opbench.py --conv tcr --out gpurun_out/opbench21.json > gpurun_out/opbench21.log 2>&1; echo "== opbench rc=$?"; grep "conv_total" gpurun_out/opbench21.log
timeout 600 python bench.py > gpurun_out/bench21.json 2> gpurun_out/bench21.err; echo "== bench"; cat gpurun_out/bench21.json | cut -c1-3500; tail -3 gpurun_out/bench21.err
