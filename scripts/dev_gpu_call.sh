cd /root/repo
timeout 900 python -m pytest tests/test_classic_gpu.py -m gpu -q > gpurun_out/s2_classic_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s2_classic_tests.log
tail -3 gpurun_out/s2_classic_tests.log
echo "== production"; timeout 300 python scripts/regime_bench.py 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:(round(v['my_per_s']/1e3), v['checksum_meanT_last']) for k,v in d.items() if isinstance(v,dict)})"
echo "== production, EBM_NO_UPAR"; EBM_NO_UPAR=1 timeout 300 python scripts/regime_bench.py 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:(round(v['my_per_s']/1e3), v['checksum_meanT_last']) for k,v in d.items() if isinstance(v,dict)})"
