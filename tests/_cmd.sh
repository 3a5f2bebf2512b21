timeout 300 python tests/run_profile.py rosen10 1000 2>&1 | tail -1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv python tests/train_bench.py profile > /dev/null 2>&1
python profiles/summarise.py launches gpurun_out/train_launches.csv
