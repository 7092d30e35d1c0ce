set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r01_gpu_tests.log
python bench.py 2>gpurun_out/bench_err.log | tee gpurun_out/r01_bench_n1.json | cut -c1-200
python bench.py --impl reference --steps 3 2>>gpurun_out/bench_err.log | tee gpurun_out/r01_bench_ref_n1.json | cut -c1-200
for w in dna178k ultralong rna40k; do python bench.py --workload $w --no-cpu --no-svbzd --steps 5 2>>gpurun_out/bench_err.log | tee gpurun_out/r01_bench_$w.json | cut -c1-120; done
python bench.py --mode event --no-cpu --no-svbzd --steps 5 2>>gpurun_out/bench_err.log | tee gpurun_out/r01_bench_event_only.json | cut -c1-120
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"walk|emit|scan_counts|verify|chunk_count|count_tile|read_event|init_reads|build_seq|gen_|sum_fixups|svb|stat_|pa_kernel|ent_kernel|jnn_" -c 600 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-siblings --reads-per-step 4096 --e2e-reads 256 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:walk_chunks -s 3 -c 1 -f -o gpurun_out/r01_walk_chunks python bench.py --steps 1 --warmup 3 --no-cpu --no-svbzd --no-siblings --reads-per-step 4096 --e2e-reads 64 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:emit_events -s 3 -c 1 -f -o gpurun_out/r01_emit_events python bench.py --steps 1 --warmup 3 --no-cpu --no-svbzd --no-siblings --reads-per-step 4096 --e2e-reads 64 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:svb_write -s 1 -c 1 -f -o gpurun_out/r01_svb_write python bench.py --steps 1 --warmup 3 --no-cpu --reads-per-step 1024 --e2e-reads 16384 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:ent_kernel -s 1 -c 1 -f -o gpurun_out/r01_ent_kernel python bench.py --steps 1 --warmup 3 --no-cpu --no-svbzd --reads-per-step 4096 --e2e-reads 64 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:jnn_walk -s 1 -c 1 -f -o gpurun_out/r01_jnn_walk python bench.py --steps 1 --warmup 3 --no-cpu --no-svbzd --reads-per-step 4096 --e2e-reads 64 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:stat_moments -s 1 -c 1 -f -o gpurun_out/r01_stat_moments python bench.py --steps 1 --warmup 3 --no-cpu --no-svbzd --reads-per-step 4096 --e2e-reads 64 > /dev/null 2>&1
python tools/fallback_cost.py 2>&1 | tail -2 | tee gpurun_out/r01_glitch_cost.jsonl
python tools/cli_bench.py --reads 8000 2>gpurun_out/cli_bench_err.log | tee gpurun_out/r01_cli_bench.jsonl | cut -c1-200
python tools/stat_time.py 2>&1 | tail -3 | tee gpurun_out/r01_stat_pa_time.jsonl
