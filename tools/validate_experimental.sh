#!/bin/bash
# First GPU call of the next round: run the engine variants that were written at the end of round 1 without hardware
# (tc_epi_groups, clf_grad_in_bwd, fused_head, tc_grouped_wgrad) through their gated equivalence test, one process per variant with a timeout of its own
# (a tcgen05 protocol bug hangs or traps the context), then time each variant that passed against the default path.
#   gpurun --timeout 900 -- 'tools/validate_experimental.sh 2>&1 | tee gpurun_out/validate_experimental.log'
cd "$(dirname "$0")/.."
export PSVAE_TEST_EXPERIMENTAL=1
pass=()
for v in "opts0" "opts1" "opts2" "opts3"; do
  echo "== test_experimental_engine_variants[$v]"
  if timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "experimental and $v" 2>&1 | tail -4; then :; fi
  rc=${PIPESTATUS[0]}
  echo "   rc=$rc"
  [ "$rc" = "0" ] && pass+=("$v")
done
echo "passed: ${pass[*]}"
get() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4), d['clocks']['sm_mhz'])"; }
declare -A OPTS=( [opts0]="--opt tc_epi_groups=1" [opts1]="--opt clf_grad_in_bwd=1" [opts2]="--opt clf_grad_in_bwd=1 --opt fused_head=1" [opts3]="--opt tc_grouped_wgrad=1" )
for v in "${pass[@]}"; do
  for i in 1 2 3; do
    echo -n "default : "; timeout 120 python bench.py --no-secondary 2>/dev/null | get
    echo -n "${OPTS[$v]} : "; timeout 120 python bench.py --no-secondary ${OPTS[$v]} 2>/dev/null | get
  done
done
