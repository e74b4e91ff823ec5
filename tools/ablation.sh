#!/bin/bash
# marginal cost of each kernel group under 8-lane overlap: throughput with the group removed (results are garbage)
for g in none expand gram stats assemble diag fused backsolve predict "diag,fused,backsolve"; do
  RG_DBG_SKIP=$g timeout 120 python bench.py --no-cpu --no-step2 --steps 5 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('skip %-22s ms/step %.2f' % ('$g', j['ms_per_step']))"
done
