(time python -m pytest tests -m gpu -x -q) > gpurun_out/r2l_pytest.log 2>&1; tail -25 gpurun_out/r2l_pytest.log
python bench.py --workload config4 --steps 10 > gpurun_out/r2l_config4.json 2> gpurun_out/r2l_config4.err; tail -2 gpurun_out/r2l_config4.err; python - <<PY
import json
d=json.load(open("gpurun_out/r2l_config4.json"))
print("config4", round(d["value"]), round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],2), "tfce", round(d["roofline"]["kernel_ms_per_launch"],2), "fit", round(d["roofline"]["fit"]["ms_per_launch"],2), round(d["roofline"]["fit"]["achieved"],1), d["cpu_baseline"])
PY
