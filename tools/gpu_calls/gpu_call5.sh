mkdir -p gpurun_out
timeout 300 python tests/tc_probe.py > gpurun_out/tcr_probe.log 2>&1; echo "== probe"; tail -16 gpurun_out/tcr_probe.log | cut -c1-200
timeout 900 python -m pytest tests/test_parity_gpu.py -q --timeout 300 -k "tcr or encoder or style_vectors" > gpurun_out/pytest_tcr.log 2>&1; echo "== tcr tests"; tail -6 gpurun_out/pytest_tcr.log | cut -c1-200
timeout 600 python tools/opbench.py --conv tcr --out gpurun_out/opbench_tcr.json > gpurun_out/opbench_tcr.log 2>&1; echo "== opbench tcr"; grep -E "modconv|conv_total" gpurun_out/opbench_tcr.log | cut -c1-200
timeout 600 python tools/opbench.py --conv tcr --unmasked --layers 'c5@32,c6^64,c7@64,c8^128,c9@128,c10^256,c11@256' --out gpurun_out/opbench_tcr_unmasked.json > gpurun_out/opbench_tcr_unmasked.log 2>&1; echo "== opbench tcr unmasked"; grep modconv gpurun_out/opbench_tcr_unmasked.log | cut -c1-200
E4S_B200_CONV=tcr timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tcr.log 2>&1; echo "== bench tcr"; tail -1 gpurun_out/bench_tcr.log | cut -c1-600; python -c "
import json;d=json.loads(open('gpurun_out/bench_tcr.log').read().strip().splitlines()[-1]);print('inversion',d.get('inversion'));print('e2e',d.get('e2e'))"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'modconv3x3' -c 5 -o gpurun_out/prof_r1_tcr python tools/ncu_targets.py --conv tcr > gpurun_out/ncu_tcr.log 2>&1; tail -2 gpurun_out/ncu_tcr.log
