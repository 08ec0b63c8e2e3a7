# round 2, call 23: re-capture of the static evidence after the tap-mask change to modconv_tcr.cu (kernel source hash)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'modconv3x3' -f -o gpurun_out/r2c23_ncu_layers python tools/opbench.py --only-conv --once --conv auto --out gpurun_out/r2c23_once.json > gpurun_out/r2c23_ncu_layers.log 2>&1; echo "== ncu layers rc=$?"; tail -2 gpurun_out/r2c23_ncu_layers.log
E4S_BENCH_PROFILE_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2c23_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-e2e --inversion-steps 0 --faceswap-pairs 0 --gpen-batch 0 > gpurun_out/r2c23_launches.log 2>&1; echo "== ncu launches rc=$?"; wc -l gpurun_out/r2c23_launches.csv
timeout 900 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'modconv3x3_tcr|instnorm|norm_residual|region_mean' -c 200 --csv --log-file gpurun_out/r2c23_encoder.csv python - > gpurun_out/r2c23_encoder.log 2>&1 <<'PY'
import sys, torch
sys.path.insert(0, '.')
from bench import build_net, face_label_maps
from e4s_b200.masks import labelMap2OneHot
dev = torch.device('cuda:0')
net = build_net(1024, 12, dev)
img = torch.randn(16, 3, 1024, 1024, generator=torch.Generator().manual_seed(3)).to(dev)
onehot = labelMap2OneHot(face_label_maps(16, 12, 'faces', 5).to(dev), 12)
with torch.no_grad():
    net.get_style_vectors(img, onehot)
torch.cuda.synchronize()
PY
echo "== ncu encoder rc=$?"; wc -l gpurun_out/r2c23_encoder.csv
