# round 2, call 22: encoder stride-2 convolutions as four taps on a space-to-depth tensor, 1x1 shortcuts as the centre tap (tap mask)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2c22_pytest.log 2>&1; echo "== pytest rc=$?"; tail -4 gpurun_out/r2c22_pytest.log
for s in 0 1; do
  E4S_B200_ENC_S2D=$s timeout 300 python tools/enc_bench.py --out gpurun_out/r2c22_enc_bench_s2d$s.json > gpurun_out/r2c22_enc_bench_s2d$s.log 2>&1; echo "== enc_bench S2D=$s rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2c22_enc_bench_s2d$s.json')); print(d['ms_per_call'], {k:(v['launches'],v['ms']) for k,v in d['entries'].items()})"
done
timeout 300 python tools/opbench.py --only-conv --conv auto --out gpurun_out/r2c22_opbench_auto.json > gpurun_out/r2c22_opbench_auto.log 2>&1; echo "== opbench auto rc=$?"; tail -1 gpurun_out/r2c22_opbench_auto.log
timeout 600 python bench.py --steps 5 --no-e2e --no-cpu-baseline --no-gpu-baseline --inversion-steps 0 --gpen-batch 16 > gpurun_out/r2c22_bench_legs.json 2> gpurun_out/r2c22_bench_legs.err; echo "== bench legs rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2c22_bench_legs.json')); print(d['value'], d['faceswap'], {k:v for k,v in d['gpen'].items() if k!='config'})"
