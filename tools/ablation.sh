#!/bin/bash
# marginal cost of each kernel group under 8-lane overlap: throughput with the group removed (results are garbage,
# parity check off).  Groups: expand gram stats assemble mx (mixed solver) predict; chol/backsolve with RG_B200_SOLVER=f64.
for g in none mxall stats predict "mxall,predict"; do
  RG_DBG_SKIP=$g timeout 160 python bench.py --no-cpu --no-step2 --steps 5 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('skip %-24s ms/step %6.2f  (%.0f SNPs/s)' % ('$g', j['ms_per_step'], j['value']))"
done
RG_B200_SOLVER=f64 timeout 160 python bench.py --no-cpu --no-step2 --steps 5 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fp64 solver, nothing skipped   ms/step %6.2f  (%.0f SNPs/s)' % (j['ms_per_step'], j['value']))"
