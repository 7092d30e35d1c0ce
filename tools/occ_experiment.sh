for sm in 0 56000 75000 110000 200000; do
SGPU_WALK_SMEM=$sm python bench.py --no-cpu --steps 5 --e2e-reads 64 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('smem $sm', d['roofline']['stage_ms_per_step']['walk_chunks'])"
done
