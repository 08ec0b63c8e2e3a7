# round 2, call 12: K-split small GEMM, noise prefetch, discriminator mirror - tests, timing, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c12_pytest.log 2>&1; echo "== pytest rc=$?"; tail -5 gpurun_out/r2c12_pytest.log
timeout 300 python - > gpurun_out/r2c12_linear_timing.log 2>&1 <<'PY'
import torch, sys
sys.path.insert(0, '.')
from e4s_b200 import kernels as K
def t(fn, n=30):
    for _ in range(3): fn()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
dev = 'cuda'
print("graph-replayed device time per call (us)")
for m, n, k in [(192, 512, 512), (192, 64, 512), (12, 512, 512), (1, 512, 512)]:
    x = torch.randn(m, k, device=dev); w = torch.randn(n, k, device=dev); b = torch.randn(n, device=dev)
    print(f"linear [{m},{k}]x[{k},{n}]: own {t(lambda: K.linear(x, w, b)):.1f}  torch F.linear {t(lambda: torch.nn.functional.linear(x, w, b)):.1f}")
    s = torch.randn(m, k, device=dev); wsq = torch.rand(n, k, device=dev)
    print(f"  demod rows {m} cin {k} cout {n}: {t(lambda: K.demod(s, wsq)):.1f}")
for g, m, n, k in [(12, 16, 512, 1280), (12, 16, 6656, 512), (12, 1, 512, 1280), (12, 1, 6656, 512)]:
    x = torch.randn(g, m, k, device=dev); w = torch.randn(g, n, k, device=dev); b = torch.randn(g, n, device=dev); wt = w.transpose(1, 2).contiguous()
    gy = torch.randn(g, m, n, device=dev)
    print(f"grouped G={g} [{m},{k}]x[{k},{n}]: own fwd {t(lambda: K.linear(x, w, b, 0.01)):.1f}  own bwd {t(lambda: K.linear(gy, w, None, 1.0, w_is_kn=True)):.1f}  torch baddbmm {t(lambda: torch.baddbmm(b.unsqueeze(1), x, wt)):.1f}  torch bwd bmm {t(lambda: torch.bmm(gy, w)):.1f}")
PY
echo "== linear timing rc=$?"; cat gpurun_out/r2c12_linear_timing.log
timeout 300 python tools/opbench.py --only-conv --conv auto --out gpurun_out/r2c12_opbench_auto.json > gpurun_out/r2c12_opbench_auto.log 2>&1; echo "== opbench auto rc=$?"; grep -o '"kernel": "[^"]*", "ms": [0-9.]*' gpurun_out/r2c12_opbench_auto.log; tail -1 gpurun_out/r2c12_opbench_auto.log
timeout 900 python bench.py --no-cpu-baseline --no-gpu-baseline --no-loss-nets --faceswap-pairs 0 --gpen-batch 0 > gpurun_out/r2c12_bench.json 2> gpurun_out/r2c12_bench.err; echo "== bench rc=$?"; cut -c1-200 gpurun_out/r2c12_bench.json; tail -3 gpurun_out/r2c12_bench.err
python - <<'PY'
import json
p=json.load(open('gpurun_out/r2c12_bench.json'))
for k,v in sorted(p['kernels'].items(), key=lambda kv:-kv[1]['ms']): print('   ',k, round(v['ms']/p['steps'],3), v['launches']//p['steps'])
inv=p['inversion']; print('inv', inv['ms_per_step'], inv['cuda_graph'], inv['batched'])
for k,v in sorted(inv['kernels'].items(), key=lambda kv:-kv[1]['ms_per_step'])[:6]: print('   ',k,v)
PY
