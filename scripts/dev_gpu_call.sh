cd /root/repo
timeout 900 python -m pytest tests/test_classic_gpu.py -m gpu -q > gpurun_out/s2_classic_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s2_classic_tests.log
tail -3 gpurun_out/s2_classic_tests.log
echo "== production"; timeout 300 python scripts/regime_bench.py 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:round(v['my_per_s']/1e3) for k,v in d.items() if isinstance(v,dict)})"
for t in g4m4 g5m3 g6m4 g7m4; do echo "== $t"; EBM_CUDA_LIB=/root/repo/energybalancemodel.jl_b200/lib/libebm_dev_$t.so timeout 300 python scripts/regime_bench.py 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:round(v['my_per_s']/1e3) for k,v in d.items() if isinstance(v,dict)})"; done
