set -x
python bench.py --workload tiny --steps 3 > gpurun_out/r2h_tiny.json 2> gpurun_out/r2h_tiny.err; tail -3 gpurun_out/r2h_tiny.err; cat gpurun_out/r2h_tiny.json
python bench.py --workload tiny --job 400 --job-check 40 > gpurun_out/r2h_tinyjob.json 2> gpurun_out/r2h_tinyjob.err; tail -3 gpurun_out/r2h_tinyjob.err; cat gpurun_out/r2h_tinyjob.json
python bench.py --workload config2_3mm --steps 6 > gpurun_out/r2h_3mm.json 2> gpurun_out/r2h_3mm.err; tail -3 gpurun_out/r2h_3mm.err; cat gpurun_out/r2h_3mm.json
