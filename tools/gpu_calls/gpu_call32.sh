mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest32.log 2>&1; echo "== pytest rc=$?"; tail -4 gpurun_out/pytest32.log
timeout 600 python bench.py > gpurun_out/bench32.json 2> gpurun_out/bench32.err; echo "== bench rc=$?"; cut -c1-700 gpurun_out/bench32.json; tail -3 gpurun_out/bench32.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches32.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --inversion-steps 0 > gpurun_out/launches32.log 2>&1; echo "== ncu launches rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'upfirdn2d_fir4|modconv3x3|torgb' -c 8 -f -o gpurun_out/prof32 python tools/ncu_targets.py --conv tcr > gpurun_out/ncu32.log 2>&1; echo "== ncu full rc=$?"; tail -2 gpurun_out/ncu32.log
timeout 200 python tools/bwd_layers.py > gpurun_out/bwd32.log 2>&1; echo "== bwd rc=$?"; tail -40 gpurun_out/bwd32.log
