# usage: TAG=name bash tools/gpu/run_tests_bench.sh   -- GPU tests + bench + ncu launch list
T=${TAG:-x}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_$T.log
timeout 400 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_$T.log 2>&1; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_$T.log').read().strip().splitlines()[-1])
    print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'], {k:(v.get('value'), v.get('tensor_frac'), v.get('error')) for k,v in d['secondary'].items()})
except Exception as e:
    print('bench parse fail', e); print(open('gpurun_out/bench_$T.log').read()[-2000:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 160 --csv --log-file gpurun_out/launches_$T.csv python bench.py --steps 2 --warmup 3 --no-secondary > gpurun_out/ncu_$T.log 2>&1; echo "ncu rc=$?"
