mkdir -p gpurun_out/r2final
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2final/launches_config4.csv python bench.py --workload config4 --steps 2 --warmup 3 --no-cpu > gpurun_out/r2final/bench_config4_under_ncu.json 2> /dev/null
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"glm_dmma_multi|glm_pack" -c 2 -o /tmp/ncu_c4 python bench.py --workload config4 --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
python scripts/ncu_summary.py /tmp/ncu_c4.ncu-rep > gpurun_out/r2final/ncu_config4_fit_summary.txt 2>&1
python scripts/ncu_hot_lines.py /tmp/ncu_c4.ncu-rep 25 > gpurun_out/r2final/ncu_config4_fit_hot_lines.txt 2>&1
python scripts/launch_summary.py gpurun_out/r2final/launches_config4.csv > gpurun_out/r2final/launches_config4_summary.txt 2>&1
wc -l gpurun_out/r2final/*.txt gpurun_out/r2final/*.csv
