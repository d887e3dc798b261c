python mptc_b200/build.py --force >/dev/null 2>&1
echo "== unroll 1 (+ early abort)"; timeout 300 python profiles/micro/k2_ab.py 2>&1 | cut -c1-150
MPTC_EXTRA_NVCC_FLAGS="-DMPTC_K2W_EVAL_UNROLL=2" python mptc_b200/build.py --force >/dev/null 2>&1
echo "== unroll 2"; timeout 300 python profiles/micro/k2_ab.py 2>&1 | cut -c1-150
python mptc_b200/build.py --force >/dev/null 2>&1
timeout 900 python -m pytest tests/test_gpu_inter_variants.py tests/test_gpu_parity_small.py -x -q -m gpu 2>&1 | tail -2
timeout 300 python bench.py --steps 4 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], {k:(round(v['value']), round(v['ms_per_step'],1)) for k,v in d['legs']['robustness'].items() if isinstance(v,dict)})"
