# round 2, call 6: loss-network numerics on the GPU (diagnostic), own small-GEMM kernel tests, encoder timing
mkdir -p gpurun_out
timeout 600 python tools/loss_diag.py > gpurun_out/r2c06_loss_diag.log 2>&1; echo "== loss diag rc=$?"; cat gpurun_out/r2c06_loss_diag.log | cut -c1-330
timeout 900 python -m pytest tests -m gpu -q -k "linear or local_mlps or net3 or generator or inversion or graphed or styled or swap" > gpurun_out/r2c06_pytest.log 2>&1; echo "== pytest rc=$?"; tail -8 gpurun_out/r2c06_pytest.log
timeout 300 python tools/enc_bench.py --out gpurun_out/r2c06_enc_bench.json > gpurun_out/r2c06_enc_bench.log 2>&1; echo "== enc bench rc=$?"; tail -2 gpurun_out/r2c06_enc_bench.log | cut -c1-1500
timeout 600 python bench.py --no-cpu-baseline --no-gpu-baseline --no-loss-nets --faceswap-pairs 0 --gpen-batch 0 --inversion-batch 0 > gpurun_out/r2c06_bench.json 2> gpurun_out/r2c06_bench.err; echo "== bench rc=$?"; cut -c1-300 gpurun_out/r2c06_bench.json; tail -2 gpurun_out/r2c06_bench.err
