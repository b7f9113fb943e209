# Round measurement recipe (run under gpurun on one B200): ncu --set full capture of the traversal kernel, launch list, bench lines of every
# workload and of the reference arm into gpurun_out/; scripts/ncu_summary.py turns the captures into profiles/*.md.
set -x
ncu --set full --clock-control none --import-source on -k regex:kb_traverse -s 2 -c 1 -o gpurun_out/prof_traverse_r01_v8 -f python bench.py --steps 2 --warmup 1 > gpurun_out/b_ncu8.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_v8.csv python bench.py --steps 2 --warmup 1 > gpurun_out/b_ncu8b.log 2>&1
python bench.py > gpurun_out/bench_r01_v8.json 2> gpurun_out/bench_r01_v8.err
python bench.py --impl reference > gpurun_out/bench_r01_v8_ref.json 2> gpurun_out/bench_r01_v8_ref.err
for w in c1 c3 c4 c5; do python bench.py --workload $w > gpurun_out/bench_r01_v8_$w.json 2> gpurun_out/bench_r01_v8_$w.err; done
tail -c 600 gpurun_out/bench_r01_v8.json; for w in c1 c3 c4 c5; do head -c 300 gpurun_out/bench_r01_v8_$w.json; echo; done
