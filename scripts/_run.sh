O=gpurun_out
TAG=r04l
NCU="ncu --set full --clock-control none --import-source on -f"
python scripts/one_search.py cfg2 0:0 3 > $O/${TAG}_search_cfg2_plain.log 2>&1 || { echo plain failed; exit 1; }
$NCU -k regex:rvq_search -s 2 -c 1 -o $O/${TAG}_search_cfg2 python scripts/one_search.py cfg2 0:0 3 > $O/${TAG}_search_cfg2_ncu.log 2>&1
tail -2 $O/${TAG}_search_cfg2_ncu.log
ncu -i $O/${TAG}_search_cfg2.ncu-rep --page raw --csv > $O/${TAG}_search_cfg2_raw.csv 2>/dev/null
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary --no-module --sustained-s 0 --e2e-steps 1 > $O/${TAG}_bench_short.json 2>$O/${TAG}_bench_short.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary --no-module --sustained-s 0 --e2e-steps 1 > $O/${TAG}_ncu_bench.log 2>&1
grep -c rvq_search $O/${TAG}_launches.csv
