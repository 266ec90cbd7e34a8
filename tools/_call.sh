mkdir -p gpurun_out/r2
python -m pytest tests/test_gpu_trainloop.py -m gpu -q -x 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline 2>&1 | tail -1 > gpurun_out/r2/bench30.json; python - <<'PY'
import json; d=json.loads(open('gpurun_out/r2/bench30.json').read()); print(d['ms_per_step'], d['value'], d['e2e'], d['roofline']['frac'], d['roofline'].get('launch_us'), d['roofline'].get('single_launch_us'), d['gpu_launches'])
PY
