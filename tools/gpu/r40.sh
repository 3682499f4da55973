set -x
python build/bm_probe.py; PSVAE_B200_LIB=$PWD/build/ab/libpsvae_fastbm.so python build/bm_probe.py
for lib in "" build/ab/libpsvae_fastbm.so; do
  export PSVAE_B200_LIB=${lib:+$PWD/$lib}; [ -z "$lib" ] && unset PSVAE_B200_LIB
  python tools/langevin_bench.py 1048576; python tools/philox_bench.py 2>&1 | tail -2
  timeout 200 python bench.py --steps 100 --warmup 20 --no-secondary > gpurun_out/r40_tmp.json 2>/dev/null; python -c "
import json;d=json.loads(open('gpurun_out/r40_tmp.json').read().strip().splitlines()[-1]);print('bench',d['ms_per_step'])"
  python tools/sample_bench.py 2>&1 | grep "chain=1 N=125"
done
