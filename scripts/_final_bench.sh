#!/bin/bash
# Final bench lines of every BASELINE configuration (one GPU), written to gpurun_out/r2final/.
mkdir -p gpurun_out/r2final
for w in config2 config2_3mm config1 config3 config4 config5; do
  timeout 500 python bench.py --workload $w 2>/dev/null | tail -1 > gpurun_out/r2final/bench_$w.json
  python - <<PY
import json
d = json.load(open("gpurun_out/r2final/bench_$w.json"))
print("$w", round(d["value"], 1), round(d["e2e"]["value"], 1), "tfce ms", round(d["roofline"]["kernel_ms_per_launch"], 2), "frac", round(d["roofline"]["frac"], 3),
      "fit ms", round(d["roofline"]["fit"]["ms_per_launch"], 2), "rows identical:", d["cpu_baseline"]["sample"][-5:])
PY
done
