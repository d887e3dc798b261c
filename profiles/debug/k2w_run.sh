set -x
timeout 900 python -m pytest tests/test_gpu_inter_variants.py tests/test_gpu_parity_small.py tests/test_gpu_full_golden.py tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | tail -5
timeout 300 python profiles/micro/k2_ab.py 2>&1 | cut -c1-170
for cfg in "16 0" "8 50" "16 20"; do set -- $cfg; SA=$1 THR=$2 timeout 300 python profiles/micro/k2_ab.py | cut -c1-170; done
