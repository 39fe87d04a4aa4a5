python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --job 10000 --job-check 32 > gpurun_out/r2p_job2.json 2> gpurun_out/r2p_job2.err; tail -3 gpurun_out/r2p_job2.err; python -c "
import json; d=json.load(open('gpurun_out/r2p_job2.json')); print(d['value'], d['breakdown_s'], d['parity'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 > gpurun_out/r2p_bench2.json 2> gpurun_out/r2p_bench2.err; tail -3 gpurun_out/r2p_bench2.err; python -c "
import json; d=json.load(open('gpurun_out/r2p_bench2.json')); print('N=2', round(d['value']), round(d['e2e']['value']), d['config']['gather'])"
python -m pytest tests/test_gpu_reference_drivers.py -q -m gpu -k step2 2>&1 | tail -3
