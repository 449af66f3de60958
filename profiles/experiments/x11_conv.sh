O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_dropin_gpu.py -m gpu -q -x -k "head_conv or detect_convs or headconv" > $O/x11_pytest.log 2>&1; tail -3 $O/x11_pytest.log
timeout 200 python profiles/experiments/check_detconv.py 2>&1 | tail -12
