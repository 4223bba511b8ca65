cd /root/repo
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2b_gpu_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2b_gpu_tests.log
tail -5 gpurun_out/r2b_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
