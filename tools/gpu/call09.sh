# round 2, call 9: small-GEMM kernel v2 (modulations, LocalMLPs, demodulation) timing + tests + bench
mkdir -p gpurun_out
timeout 300 python - > gpurun_out/r2c09_linear_timing.log 2>&1 <<'PY'
import torch, sys
sys.path.insert(0, '.')
from e4s_b200 import kernels as K
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
dev = 'cuda'
for m, n, k in [(192, 512, 512), (192, 64, 512), (16, 512, 512), (12, 512, 512), (1, 512, 512)]:
    x = torch.randn(m, k, device=dev); w = torch.randn(n, k, device=dev); b = torch.randn(n, device=dev)
    print(f"linear [{m},{k}]x[{k},{n}]: own {t(lambda: K.linear(x, w, b)):.1f} us   torch F.linear {t(lambda: torch.nn.functional.linear(x, w, b)):.1f} us")
    s = torch.randn(m, k, device=dev); wsq = torch.rand(n, k, device=dev)
    print(f"  demod rows {m} cin {k} cout {n}: {t(lambda: K.demod(s, wsq)):.1f} us")
for g, m, n, k in [(12, 16, 512, 1280), (12, 16, 6656, 512), (12, 1, 512, 1280), (12, 1, 6656, 512)]:
    x = torch.randn(g, m, k, device=dev); w = torch.randn(g, n, k, device=dev); b = torch.randn(g, n, device=dev)
    print(f"grouped linear G={g} [{m},{k}]x[{k},{n}]: own {t(lambda: K.linear(x, w, b, 0.01)):.1f} us   torch baddbmm {t(lambda: torch.baddbmm(b.unsqueeze(1), x, w.transpose(1, 2))):.1f} us")
PY
echo "== linear timing rc=$?"; cat gpurun_out/r2c09_linear_timing.log
timeout 600 python -m pytest tests -m gpu -q -k "linear or local_mlps or demod or net3 or generator_golden or inversion_loop or graphed or encoder or rgi or style_vectors" > gpurun_out/r2c09_pytest.log 2>&1; echo "== pytest rc=$?"; tail -4 gpurun_out/r2c09_pytest.log
timeout 900 python bench.py --no-cpu-baseline --no-gpu-baseline --no-loss-nets --faceswap-pairs 0 --gpen-batch 0 > gpurun_out/r2c09_bench.json 2> gpurun_out/r2c09_bench.err; echo "== bench rc=$?"; cut -c1-300 gpurun_out/r2c09_bench.json; tail -3 gpurun_out/r2c09_bench.err
python - <<'PY'
import json
p=json.load(open('gpurun_out/r2c09_bench.json'))
for k,v in sorted(p['kernels'].items(), key=lambda kv:-kv[1]['ms']): print('   ',k, round(v['ms']/p['steps'],3), v['launches']//p['steps'])
inv=p['inversion']; print('inv', inv['ms_per_step'], inv['cuda_graph'], inv['batched'])
PY
