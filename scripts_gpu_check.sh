#!/bin/bash
# one GPU-box call: smoke, GPU tests, bench (+ full per-kernel table); logs under gpurun_out/
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -q --tb=short --timeout 900 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
timeout 900 python bench.py ${BENCH_ARGS:---steps 10 --warmup 3} > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -3 gpurun_out/smoke.log; tail -25 gpurun_out/pytest_gpu.log
python - <<'PY'
import json
try:
    l=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
    print('ms/step', l['ms_per_step'], 'value', l['value'], 'e2e', l['e2e']['value'], 'launches', l['gpu_launches'], 'clocks', l['clocks'])
    for k in l['roofline']['kernels']: print('  %-42s %8.4f ms  share %.3f  GB/s %s' % (k['kernel'], k['ms_per_step'], k['share'], k['achieved_gbs'] and round(k['achieved_gbs'])))
    print('step_frac', l['roofline'].get('step_frac'), 'extras', l['extras'], 'cpu', l['cpu_baseline'] and l['cpu_baseline']['value'])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench.err').read()[-2000:])
PY
