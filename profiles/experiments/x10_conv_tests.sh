O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_dropin_gpu.py -m gpu -q -x -k "head_conv or detect_convs or headconv" > $O/x10_pytest.log 2>&1; tail -15 $O/x10_pytest.log
