mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest33.log 2>&1; echo "== pytest rc=$?"; tail -5 gpurun_out/pytest33.log
timeout 200 python tools/bwd_layers.py > gpurun_out/bwd33_default.log 2>&1; echo "== bwd default rc=$?"; grep -v tcr_fwd gpurun_out/bwd33_default.log | tail -20
E4S_B200_DGRAD_SPLIT=1,1 timeout 200 python tools/bwd_layers.py > gpurun_out/bwd33_nosplit.log 2>&1; echo "== bwd nosplit rc=$?"; tail -1 gpurun_out/bwd33_nosplit.log
E4S_B200_DGRAD_SPLIT=1,1 E4S_B200_NTILE=256 timeout 200 python tools/bwd_layers.py > gpurun_out/bwd33_old.log 2>&1; echo "== bwd old rc=$?"; tail -1 gpurun_out/bwd33_old.log
timeout 600 python bench.py > gpurun_out/bench33.json 2> gpurun_out/bench33.err; echo "== bench rc=$?"; cut -c1-400 gpurun_out/bench33.json; tail -3 gpurun_out/bench33.err
timeout 300 python tools/opbench.py --only-hbm --out gpurun_out/opbench33_hbm.json > gpurun_out/opbench33_hbm.log 2>&1; echo "== opbench hbm rc=$?"; cat gpurun_out/opbench33_hbm.log | cut -c1-200
E4S_BENCH_PROFILE_RANGE=1 timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches33.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --inversion-steps 0 > gpurun_out/launches33.log 2>&1; echo "== ncu launches rc=$?"; wc -l gpurun_out/launches33.csv
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'modconv3x3_tcr' -f -o gpurun_out/prof33_layers python tools/opbench.py --only-conv --once --conv tcr --out gpurun_out/opbench33_once.json > gpurun_out/ncu33.log 2>&1; echo "== ncu full rc=$?"; tail -2 gpurun_out/ncu33.log
