mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'modconv3x3_tcr' -c 5 -f -o gpurun_out/prof_r1_tcr_3mma python tools/ncu_targets.py --conv tcr > gpurun_out/ncu_tcr_3mma.log 2>&1; echo "== ncu rc=$?"; tail -3 gpurun_out/ncu_tcr_3mma.log; ls -la gpurun_out/*.ncu-rep
