set -x
timeout 900 python -m pytest tests/test_gpu_inter_variants.py tests/test_gpu_parity_small.py tests/test_gpu_full_golden.py tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | tail -5
for p in 0 2; do MPTC_PRIO=$p timeout 300 python profiles/micro/k2_ab.py 2>&1 | cut -c1-170; done
MPTC_PRIO=0 MPTC_K2_NO_INT8=1 timeout 300 python profiles/micro/k2_ab.py 2>&1 | cut -c1-170
for cfg in "16 0" "8 50"; do set -- $cfg; SA=$1 THR=$2 MPTC_PRIO=0 timeout 300 python profiles/micro/k2_ab.py | cut -c1-170; done
MPTC_EXTRA_NVCC_FLAGS=-DMPTC_K2_PHASE_TIMING python mptc_b200/build.py --force >/dev/null 2>&1; timeout 200 python profiles/k2_timing.py 2>&1 | tee gpurun_out/k2w_trace2.txt
