for g in none mxgemm mxpotrf mxtri mxres "mxtri,mxres" assemble; do
  RG_DBG_SKIP=$g timeout 160 python bench.py --no-cpu --no-step2 --steps 5 2>gpurun_out/abl_err.txt | python -c "
import json,sys
t=sys.stdin.read().strip().splitlines()
if t:
    j=json.loads(t[-1]); print('skip %-24s ms/step %6.2f  (%.0f SNPs/s)' % ('$g', j['ms_per_step'], j['value']))
else:
    print('skip %-24s failed: ' % '$g', open('gpurun_out/abl_err.txt').read()[-200:])"
done
