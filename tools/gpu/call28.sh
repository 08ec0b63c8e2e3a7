# round 2, call 28 (gpurun --gpus 2): the default line under torchrun on the final commit
mkdir -p gpurun_out
s0=$(date +%s)
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2c28_bench_n2.json 2> gpurun_out/r2c28_bench_n2.err; echo "== N=2 bench rc=$? in $(( $(date +%s) - s0 )) s"; cut -c1-200 gpurun_out/r2c28_bench_n2.json; tail -2 gpurun_out/r2c28_bench_n2.err
python -c "
import json; d=json.load(open('gpurun_out/r2c28_bench_n2.json')); print(d['leg_seconds']); print(d['value'], d['e2e']['value'], d['gather']['ms_alone'], d['gather']['overlap_cost_ms'], d['faceswap']['pairs_per_sec'])"
