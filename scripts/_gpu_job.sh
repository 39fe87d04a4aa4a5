(python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_wide.py tests/test_gpu_weighted.py -x -q -m gpu) 2>&1 | tail -3
python bench.py --workload config5 --steps 4 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('config5', round(d['value'],1), round(d['e2e']['value'],1), 'tfce', round(d['roofline']['kernel_ms_per_launch'],3), 'fit', round(d['roofline']['fit']['ms_per_launch'],3))"
python bench.py --workload config2_3mm --steps 6 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('config2_3mm', round(d['value'],1), round(d['e2e']['value'],1), 'tfce', round(d['roofline']['kernel_ms_per_launch'],3), 'fit', round(d['roofline']['fit']['ms_per_launch'],3))"
