for w in 4 5 6 7 8; do
VAEL_AD2_FWD_WARPS=$w VAEL_AD2_BWD_WARPS=$w python bench.py --steps 3 --warmup 3 --no-train --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); m=d['secondary']['mario_circuits']['worlds_query']; k=d['roofline']['kernels']
print('warps=$w mario fwd %.4f (%.2f) bwd %.4f (%.2f) | 10x10 fwd %.4f bwd %.4f' % (m['fwd']['ms'], m['fwd']['frac'], m['bwd']['ms'], m['bwd']['frac'], k[0]['ms'], k[1]['ms']))"
done
