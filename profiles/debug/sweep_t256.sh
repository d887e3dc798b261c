for S in 2 3 4 6; do for WR in 0; do echo "threads 256 split $S"; MPTC_LIB=$PWD/profiles/debug/variants/lib_t256.so MPTC_ROW_SPLIT=$S python profiles/rows_timing.py; done; done
echo "threads 512 split 3"; python profiles/rows_timing.py
MPTC_LIB=$PWD/profiles/debug/variants/lib_t256.so python -m pytest tests/test_gpu_parity_small.py tests/test_gpu_full_golden.py -x -q -m gpu 2>&1 | tail -2
