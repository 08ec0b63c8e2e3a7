mkdir -p gpurun_out
timeout 300 python tests/tc_probe.py > gpurun_out/probe11.log 2>&1; echo "== probe"; tail -8 gpurun_out/probe11.log | cut -c1-160
timeout 300 python tools/enc_diag.py tcp tcr > gpurun_out/enc_diag.log 2>&1; echo "== enc diag"; cat gpurun_out/enc_diag.log | cut -c1-120
