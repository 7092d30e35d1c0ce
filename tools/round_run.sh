# on the GPU box (one GPU): the measurements and captures of a round, written to gpurun_out/ (copied to profiles/ by hand)
#   gpurun --timeout 1500 -- 'bash tools/round_run.sh r02'
R=${1:-r02}
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/${R}_gpu_tests.log
python bench.py 2>gpurun_out/bench_err.log | tee gpurun_out/${R}_bench_n1.json | cut -c1-200
python bench.py --impl reference --steps 3 2>>gpurun_out/bench_err.log | tee gpurun_out/${R}_bench_ref_n1.json | cut -c1-200
for w in dna178k ultralong rna40k real; do python bench.py --workload $w --no-cpu --no-svbzd --no-others --steps 5 2>>gpurun_out/bench_err.log | tee gpurun_out/${R}_bench_$w.json | cut -c1-120; done
python bench.py --mode event --no-cpu --no-svbzd --no-others --steps 5 2>>gpurun_out/bench_err.log | tee gpurun_out/${R}_bench_event_only.json | cut -c1-120
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"walk|long_jobs|emit|scan_counts|verify|chunk_count|count_tile|read_event|init_reads|build_seq|gen_|sum_fixups|svb|stat_|pa_kernel|ent_kernel|jnn_|prefix_" -c 600 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-siblings --no-others --reads-per-step 4096 --e2e-reads 256 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:walk_chunks -s 3 -c 1 -f -o gpurun_out/${R}_walk_chunks python bench.py --steps 1 --warmup 3 --no-cpu --no-svbzd --no-siblings --no-others --reads-per-step 4096 --e2e-reads 64 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:emit_events -s 3 -c 1 -f -o gpurun_out/${R}_emit_events python bench.py --steps 1 --warmup 3 --no-cpu --no-svbzd --no-siblings --no-others --reads-per-step 4096 --e2e-reads 64 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:stat_moments -s 1 -c 1 -f -o gpurun_out/${R}_stat_moments python bench.py --steps 1 --warmup 3 --no-cpu --no-svbzd --no-others --reads-per-step 4096 --e2e-reads 64 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:stat_median -s 1 -c 1 -f -o gpurun_out/${R}_stat_median python bench.py --steps 1 --warmup 3 --no-cpu --no-svbzd --no-others --reads-per-step 4096 --e2e-reads 64 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:jnn_walk -s 1 -c 1 -f -o gpurun_out/${R}_jnn_walk python bench.py --steps 1 --warmup 3 --no-cpu --no-svbzd --no-others --reads-per-step 4096 --e2e-reads 64 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:ent_kernel -s 1 -c 1 -f -o gpurun_out/${R}_ent_kernel python bench.py --steps 1 --warmup 3 --no-cpu --no-svbzd --no-others --reads-per-step 4096 --e2e-reads 64 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:long_jobs -s 3 -c 1 -f -o gpurun_out/${R}_long_jobs python bench.py --steps 1 --warmup 3 --no-cpu --no-svbzd --no-siblings --no-others --reads-per-step 4096 --e2e-reads 64 > /dev/null 2>&1
python tools/cli_bench.py --reads 8000 --modes event-c,event,stat,pa,jnn,ent,prefix 2>gpurun_out/cli_bench_err.log | tee gpurun_out/${R}_cli_bench.jsonl | cut -c1-200
python tools/stat_time.py 2>&1 | tee gpurun_out/${R}_stat_cta_sweep.jsonl | tail -3
python tools/rna_jobs_time.py 2>&1 | tee gpurun_out/${R}_rna_jobs_time.jsonl
