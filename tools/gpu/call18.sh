# round 2, call 18: mixed-tile transform with per-tile row offsets and tap-validity masks (213 instructions per tap and thread)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2c18_pytest.log 2>&1; echo "== pytest rc=$?"; tail -4 gpurun_out/r2c18_pytest.log
timeout 300 python tools/opbench.py --only-conv --conv auto --out gpurun_out/r2c18_opbench_auto.json > gpurun_out/r2c18_opbench_auto.log 2>&1; echo "== opbench auto rc=$?"; grep -o '"kernel": "[^"]*", "ms": [0-9.]*' gpurun_out/r2c18_opbench_auto.log; tail -1 gpurun_out/r2c18_opbench_auto.log
timeout 600 python bench.py --no-cpu-baseline --no-gpu-baseline --inversion-steps 0 --faceswap-pairs 0 --gpen-batch 0 > gpurun_out/r2c18_bench.json 2> gpurun_out/r2c18_bench.err; echo "== bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c18_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','eager')}); print(d['e2e']); print(d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['share_of_step']); print(d['kernels'])
PY
E4S_B200_LIB=$PWD/e4s_b200/libe4s_b200_prof.so timeout 300 python tools/opbench.py --only-conv --conv tcr --prof --layers "c7@64,c9@128,c8^128" --out gpurun_out/r2c18_prof_tcr.json > gpurun_out/r2c18_prof_tcr.log 2>&1; echo "== prof tcr rc=$?"; grep -v '^{' gpurun_out/r2c18_prof_tcr.log | cut -c1-200
