mkdir -p gpurun_out/r2final
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2final/launches_config2.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2final/bench_config2_under_ncu.json 2> /dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2final/launches_config2_3mm.csv python bench.py --workload config2_3mm --steps 2 --warmup 3 --no-cpu > gpurun_out/r2final/bench_config2_3mm_under_ncu.json 2> /dev/null
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"pipe_|glm_dmma" -c 7 -o /tmp/ncu_c2 python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
python scripts/ncu_summary.py /tmp/ncu_c2.ncu-rep > gpurun_out/r2final/ncu_config2_summary.txt 2>&1
python scripts/ncu_hot_lines.py /tmp/ncu_c2.ncu-rep 25 > gpurun_out/r2final/ncu_config2_hot_lines.txt 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"pipe_|glm_dmma" -c 7 -o /tmp/ncu_c3 python bench.py --workload config2_3mm --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
python scripts/ncu_summary.py /tmp/ncu_c3.ncu-rep > gpurun_out/r2final/ncu_config2_3mm_summary.txt 2>&1
wc -l gpurun_out/r2final/*.txt
