# round 2, call 10: A/B of the fast issue path and the packed epilogue on the plain layers
mkdir -p gpurun_out
L="c7@64,c9@128,c11@256,c13@512,c15@1024"
for v in "" nofast scalarepi both; do
  lib=$PWD/e4s_b200/libe4s_b200${v:+_$v}.so
  E4S_B200_LIB=$lib timeout 300 python tools/opbench.py --only-conv --conv tcr --layers $L --out gpurun_out/r2c10_opbench_${v:-default}.json > gpurun_out/r2c10_opbench_${v:-default}.log 2>&1; echo "== variant '${v:-default}' rc=$?"; grep -o '"kernel": "[^"]*", "ms": [0-9.]*' gpurun_out/r2c10_opbench_${v:-default}.log
done
timeout 300 python -m pytest tests -m gpu -q -k "demod or deterministic" > gpurun_out/r2c10_pytest.log 2>&1; echo "== pytest rc=$?"; tail -3 gpurun_out/r2c10_pytest.log
