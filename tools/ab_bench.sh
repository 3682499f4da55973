#!/bin/bash
# A/B timing of two builds of the library on the same box: alternates `bench.py --no-secondary` between $1 (old .so) and the in-tree build.
# usage: tools/ab_bench.sh build/ab/libpsvae_old.so [pairs]
OLD=$1; N=${2:-3}
get() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4), d['clocks']['sm_mhz'])"; }
for i in $(seq $N); do
  echo -n "old: "; PSVAE_B200_LIB=$OLD python bench.py --no-secondary 2>/dev/null | get
  echo -n "new: "; python bench.py --no-secondary 2>/dev/null | get
done
