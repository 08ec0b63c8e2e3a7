mkdir -p gpurun_out
timeout 300 python tests/tc_probe.py > gpurun_out/tcr_parts_probe.log 2>&1; echo "== probe"; tail -16 gpurun_out/tcr_parts_probe.log | cut -c1-160
timeout 900 python -m pytest tests/test_parity_gpu.py -q --timeout 300 -k "tcr or encoder or style_vectors" > gpurun_out/pytest_tcr.log 2>&1; echo "== tcr tests"; tail -5 gpurun_out/pytest_tcr.log | cut -c1-200
timeout 900 python -m pytest tests/test_backward_gpu.py -q --timeout 300 -k "dgrad_tc" > gpurun_out/pytest_dgrad.log 2>&1; echo "== dgrad tc tests"; tail -12 gpurun_out/pytest_dgrad.log | cut -c1-200
timeout 600 python tools/opbench.py --conv tcr --out gpurun_out/opbench_tcr_parts.json > gpurun_out/opbench_tcr_parts.log 2>&1; echo "== opbench tcr"; grep -E "modconv|conv_total" gpurun_out/opbench_tcr_parts.log | cut -c1-200
E4S_B200_CONV=tcr timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --inversion-steps 0 > gpurun_out/bench_tcr_parts.log 2>&1; echo "== bench tcr"; tail -1 gpurun_out/bench_tcr_parts.log | cut -c1-400
E4S_B200_CONV=tcr E4S_B200_BWD=tc timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --inversion-steps 20 > gpurun_out/bench_inv_tc.log 2>&1; echo "== inversion tc"; python -c "
import json;d=json.loads(open('gpurun_out/bench_inv_tc.log').read().strip().splitlines()[-1]);print('inversion',d.get('inversion'))"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'modconv3x3' -c 5 -o gpurun_out/prof_r1_tcr_parts python tools/ncu_targets.py --conv tcr > gpurun_out/ncu_tcr_parts.log 2>&1; tail -2 gpurun_out/ncu_tcr_parts.log
