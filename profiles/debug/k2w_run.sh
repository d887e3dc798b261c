set -x
timeout 900 python -m pytest tests/test_gpu_inter_variants.py tests/test_gpu_parity_small.py tests/test_gpu_full_golden.py tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | tail -5
for cfg in "16 50" "16 0" "8 50"; do set -- $cfg; for k in tiled default; do SA=$1 THR=$2 MPTC_K2=$k timeout 300 python profiles/micro/k2_ab.py; done; done 2>&1 | grep -v "^+" | tee gpurun_out/k2w_ab3.txt
timeout 600 python bench.py --no-extra-legs 2>&1 | tail -1 | tee gpurun_out/k2w_bench.json
