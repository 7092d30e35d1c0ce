# development: walker time on the bench batch for several chunk lengths / warm-ups (bench.py --chunk-len / --detector-warmup)
for cl in ${CHUNKS:-480 544 736 992 1056 1120}; do for w in ${WARMUPS:-48}; do
python bench.py --no-others --no-svbzd --no-cpu --no-siblings --steps 5 --e2e-reads 64 --chunk-len $cl --detector-warmup $w "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); st=d['roofline']['stage_ms_per_step']
print('L',$cl,'W',$w, round(d['value'],1), 'walk',round(st['walk_chunks'],3),'emit',round(st['emit_events'],3),'seq',d['config']['sequential_order_reads'],'fix',d['config']['detector_fixups'])"
done; done
