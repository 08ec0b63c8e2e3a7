mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest35.log 2>&1; echo "== pytest rc=$?"; tail -4 gpurun_out/pytest35.log | cut -c1-300
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke35.log 2>&1; echo "== smoke rc=$?"; tail -2 gpurun_out/smoke35.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'dgrad_tc' -c 17 -f -o gpurun_out/prof35_dgrad python tools/bwd_layers.py > gpurun_out/ncu35.log 2>&1; echo "== ncu dgrad rc=$?"; tail -2 gpurun_out/ncu35.log
