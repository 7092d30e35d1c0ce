#!/bin/bash
# on the GPU box: time every build/variants/*.so with a short device-resident bench (restores the product .so afterwards)
cd "$(dirname "$0")/.."
cp sigtk_b200/libsigtk_b200.so /tmp/orig.so
for v in build/variants/*.so; do
  cp $v sigtk_b200/libsigtk_b200.so
  python bench.py --no-cpu --steps 6 --warmup 3 --e2e-reads 64 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['roofline']['stage_ms_per_step']
print('$v', round(d['value'],2), 'Gs/s walk', round(s['walk_chunks'],3), 'emit', round(s['emit_events'],3))"
done
cp /tmp/orig.so sigtk_b200/libsigtk_b200.so
