cd /root/repo
for t in m4 m3 m5 m7; do echo "== $t"; EBM_CUDA_LIB=/root/repo/energybalancemodel.jl_b200/lib/libebm_dev_$t.so timeout 300 python scripts/regime_bench.py --regimes partial,c4 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:(round(v['my_per_s']/1e3), v['checksum_meanT_last']) for k,v in d.items() if isinstance(v,dict)})"; done
