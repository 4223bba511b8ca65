cd /root/repo
timeout 1500 python -m pytest tests/test_classic_gpu.py -m gpu -q -k "grids or debug" > gpurun_out/s2_all_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s2_all_tests.log
tail -5 gpurun_out/s2_all_tests.log
