# round 2, call 19: evidence on the final kernels - ncu --set full of the 17 conv launches, launch list of the bench timed region
# (one CUDA-graph replay per step), ncu of the non-conv kernels of one eager forward, the default bench line and the reference arm
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'modconv3x3' -f -o gpurun_out/r2c19_ncu_layers python tools/opbench.py --only-conv --once --conv auto --out gpurun_out/r2c19_once.json > gpurun_out/r2c19_ncu_layers.log 2>&1; echo "== ncu layers rc=$?"; tail -2 gpurun_out/r2c19_ncu_layers.log
E4S_BENCH_PROFILE_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2c19_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-e2e --inversion-steps 0 --faceswap-pairs 0 --gpen-batch 0 > gpurun_out/r2c19_launches.log 2>&1; echo "== ncu launches rc=$?"; wc -l gpurun_out/r2c19_launches.csv
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:'torgb|linear_multi|label_resize|distribution_elementwise' -c 100 --csv --log-file gpurun_out/r2c19_nonconv.csv python - > gpurun_out/r2c19_nonconv.log 2>&1 <<'PY'
import sys, torch
sys.path.insert(0, '.')
from bench import build_net, face_label_maps
from e4s_b200.stylegan2.modconv import LabelPyramid
dev = torch.device('cuda:0')
net = build_net(1024, 12, dev)
codes = torch.randn(16, 12, 18, 512, generator=torch.Generator().manual_seed(100)).to(dev)
labels = face_label_maps(16, 12, 'faces', 200).to(dev)
with torch.no_grad():
    net.gen_img(None, codes, LabelPyramid(labels[:, 0], 12))
torch.cuda.synchronize()
PY
echo "== ncu non-conv rc=$?"; wc -l gpurun_out/r2c19_nonconv.csv
timeout 1500 python bench.py > gpurun_out/r2c19_bench_full.json 2> gpurun_out/r2c19_bench_full.err; echo "== full bench rc=$?"; cut -c1-600 gpurun_out/r2c19_bench_full.json
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r2c19_bench_reference.json 2> gpurun_out/r2c19_bench_reference.err; echo "== reference arm rc=$?"; cut -c1-400 gpurun_out/r2c19_bench_reference.json
