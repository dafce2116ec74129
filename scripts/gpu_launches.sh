cd /root/repo
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/launches.csv | cut -c1-200
