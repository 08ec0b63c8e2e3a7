# round 2, call 15: batched modulations (two launches per forward), forward as one CUDA graph, STK only at N = 32
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2c15_pytest.log 2>&1; echo "== pytest rc=$?"; tail -8 gpurun_out/r2c15_pytest.log
timeout 600 python bench.py --no-cpu-baseline --no-gpu-baseline --inversion-steps 0 --faceswap-pairs 0 --gpen-batch 0 > gpurun_out/r2c15_bench.json 2> gpurun_out/r2c15_bench.err; echo "== bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c15_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','eager')}); print(d['e2e']); print(d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['share_of_step']); print(d['kernels'])
PY
E4S_B200_STYLE_BATCH=0 timeout 600 python bench.py --eager --no-e2e --no-cpu-baseline --no-gpu-baseline --inversion-steps 0 --faceswap-pairs 0 --gpen-batch 0 > gpurun_out/r2c15_bench_eager_nobatch.json 2> gpurun_out/r2c15_bench_eager_nobatch.err; echo "== bench eager, per-layer modulations rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c15_bench_eager_nobatch.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','eager')})
PY
timeout 600 python bench.py --steps 5 --no-e2e --no-cpu-baseline --no-gpu-baseline --faceswap-pairs 16 --gpen-batch 16 > gpurun_out/r2c15_bench_legs.json 2> gpurun_out/r2c15_bench_legs.err; echo "== bench legs rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c15_bench_legs.json'))
inv=d['inversion']; print('inversion', {k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if 'ms' in kk}) for k,v in inv.items() if k!='kernels'})
print('faceswap', d['faceswap']); print('gpen', {k:v for k,v in d['gpen'].items() if k!='config'})
PY
