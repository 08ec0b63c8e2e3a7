mkdir -p gpurun_out
timeout 160 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench36_n2.json 2> gpurun_out/bench36_n2.err; echo "== bench n2 rc=$?"; cut -c1-260 gpurun_out/bench36_n2.json; tail -3 gpurun_out/bench36_n2.err | cut -c1-300
