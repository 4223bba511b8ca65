cd /root/repo
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 2 --warmup 1 --years 50 --miz-years 2 --no-cpu > gpurun_out/r2b_bench_8gpu_50y.json 2> gpurun_out/r2b_bench_8gpu_50y.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2b_bench_8gpu_50y.json').read().strip().splitlines()[-1]); print(d['value'], d['n_gpus'], d['ms_per_step'], d['kernel_ms_per_rank'], 'strong', d.get('strong',{}).get('value'), 'miz', d.get('miz',{}).get('value'), 'e2e', d['e2e']['value'] if d.get('e2e') else None)"
