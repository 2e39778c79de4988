O=gpurun_out/r02c
mkdir -p $O
python bench.py --no-cpu-baseline --no-e2e --no-scale-base --steps 1 --warmup 3 > $O/plain.json 2> $O/plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 30000 -c 12000 --csv --log-file $O/launches_bench_C4.csv python bench.py --no-cpu-baseline --no-e2e --no-scale-base --steps 1 --warmup 3 > $O/ncu_launches.log 2>&1
tail -2 $O/ncu_launches.log; wc -l $O/launches_bench_C4.csv
