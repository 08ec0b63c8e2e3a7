# round 2, call 5: full suite (no -x), ncu source view of the H-form kernel on c14, the new bench legs
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2c05_pytest.log 2>&1; echo "== pytest rc=$?"; tail -12 gpurun_out/r2c05_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'modconv3x3' -f -o gpurun_out/r2c05_ncu_tch python tools/opbench.py --only-conv --once --conv tch --layers "c14^1024" --out gpurun_out/r2c05_once.json > gpurun_out/r2c05_ncu_tch.log 2>&1; echo "== ncu tch rc=$?"; tail -2 gpurun_out/r2c05_ncu_tch.log
timeout 900 python bench.py > gpurun_out/r2c05_bench.json 2> gpurun_out/r2c05_bench.err; echo "== bench rc=$?"; cut -c1-300 gpurun_out/r2c05_bench.json; tail -3 gpurun_out/r2c05_bench.err
