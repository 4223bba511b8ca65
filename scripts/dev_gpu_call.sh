cd /root/repo
timeout 900 python -m pytest tests/test_miz_gpu.py -m gpu -q -s > gpurun_out/s2_miz_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s2_miz_tests.log
tail -3 gpurun_out/s2_miz_tests.log
timeout 600 python bench.py --workload miz --members 131072 --years 5 --steps 2 --warmup 1 --no-e2e --no-cpu --no-extra > gpurun_out/s2_miz_bench.json 2> gpurun_out/s2_miz_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/s2_miz_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['roofline']['frac'], d.get('nan_members'))"
