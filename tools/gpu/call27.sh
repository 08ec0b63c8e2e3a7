# round 2, call 27: the default bench line after the thread-probe trim, with per-section wall-clock
mkdir -p gpurun_out
s0=$(date +%s)
timeout 1500 python bench.py > gpurun_out/r2c27_bench_full.json 2> gpurun_out/r2c27_bench_full.err; echo "== full bench rc=$? in $(( $(date +%s) - s0 )) s"; cut -c1-200 gpurun_out/r2c27_bench_full.json
python -c "
import json; d=json.load(open('gpurun_out/r2c27_bench_full.json')); print(d['leg_seconds']); print(d['value'], d['e2e']['value'], d['cpu_baseline'])"
