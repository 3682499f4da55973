set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG:-x}.log 2>&1; echo "pytest rc=$?" 
tail -5 gpurun_out/pytest_${TAG:-x}.log
timeout 300 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_${TAG:-x}.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_${TAG:-x}.log').read().strip().splitlines()[-1])
    print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'], d['secondary'])
except Exception as e:
    print('bench parse fail', e); print(open('gpurun_out/bench_${TAG:-x}.log').read()[-2000:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/launches_${TAG:-x}.csv python bench.py --steps 2 --warmup 3 --no-secondary > gpurun_out/ncu_${TAG:-x}.log 2>&1; echo "ncu rc=$?"
