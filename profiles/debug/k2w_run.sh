timeout 900 python -m pytest tests/test_gpu_inter_variants.py tests/test_gpu_parity_small.py tests/test_gpu_full_golden.py -x -q -m gpu 2>&1 | tail -2
timeout 300 python profiles/window_sweep.py 2>&1 | tee gpurun_out/r2_window_sweep_v6.txt
