mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest34.log 2>&1; echo "== pytest rc=$?"; tail -15 gpurun_out/pytest34.log | cut -c1-300
timeout 700 python bench.py > gpurun_out/bench34.json 2> gpurun_out/bench34.err; echo "== bench rc=$?"; cut -c1-300 gpurun_out/bench34.json; tail -3 gpurun_out/bench34.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench34.json'))
    print('value',d['value'],'e2e',d['e2e']['value'],'clocks',d['clocks'])
    print('inversion',{k:d['inversion'][k] for k in ('ms_per_step','faces_per_sec_100_steps')}, d['inversion']['cuda_graph'], d['inversion']['batched'])
    print('faceswap',d['faceswap']); print('gpen',d['gpen'])
except Exception as e: print('ERR',e)
PY
timeout 400 python tools/sweep.py --out gpurun_out/sweep34.json > gpurun_out/sweep34.log 2>&1; echo "== sweep rc=$?"; cat gpurun_out/sweep34.log | cut -c1-200
timeout 200 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench34_ref.json 2> gpurun_out/bench34_ref.err; echo "== ref rc=$?"; cut -c1-400 gpurun_out/bench34_ref.json
