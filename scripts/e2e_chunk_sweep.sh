for c in 16384 32768 65536 131072 262144; do
VAEL_HOST_CHUNK_ROWS=$c python bench.py --steps 10 --warmup 3 --no-train --no-cpu --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); e=d['e2e']
print('chunk=$c e2e %.3e evals/s  %.2f ms  link frac %.3f (%.1f of %.1f GB/s)' % (e['value'], e['ms_per_step'], e['link']['frac'], e['link']['achieved_GBs_each_way'], e['link']['pinned_copy_both_ways_GBs_each_way']))"
done
