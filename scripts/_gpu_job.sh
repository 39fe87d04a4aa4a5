(python -m pytest tests/test_gpu_glm.py tests/test_glm_typeI.py tests/test_gpu_golden.py -x -q -m gpu) 2>&1 | tail -3
for t in 1 0; do TMB_GLM_TMA=$t python bench.py --steps 10 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('TMA=$t config2', round(d['value']), 'fit', round(d['roofline']['fit']['ms_per_launch'],3), round(d['roofline']['fit']['achieved'],2))"; done
for t in 1 0; do TMB_GLM_TMA=$t python bench.py --workload config4 --steps 6 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('TMA=$t config4', round(d['value']), 'fit', round(d['roofline']['fit']['ms_per_launch'],3), round(d['roofline']['fit']['achieved'],2))"; done
