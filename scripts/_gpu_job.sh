(python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_wide.py tests/test_gpu_weighted.py tests/test_gpu_tfce.py tests/test_gpu_fullsize.py tests/test_gpu_engine.py -x -q -m gpu) 2>&1 | tail -3
python bench.py --steps 20 2>/dev/null > gpurun_out/r2v_config2.json; python -c "
import json; d=json.load(open('gpurun_out/r2v_config2.json')); print('config2', round(d['value']), round(d['e2e']['value']), 'tfce', round(d['roofline']['kernel_ms_per_launch'],3), 'frac', round(d['roofline']['frac'],3), 'fit', round(d['roofline']['fit']['ms_per_launch'],3), d['cpu_baseline'])"
