python mptc_b200/build.py --force >/dev/null 2>&1
echo "== one CTA per SM (committed)"; for s in 3; do MPTC_ROW_SPLIT=$s timeout 120 python profiles/rows_timing.py; done
timeout 200 python profiles/micro/k2_ab.py | cut -c1-150
MPTC_EXTRA_NVCC_FLAGS="-DMPTC_K3R_CTAS_PER_SM=2" python mptc_b200/build.py --force >/dev/null 2>&1
echo "== two CTAs per SM (64 registers)"; for s in 2 3 4 6; do MPTC_ROW_SPLIT=$s timeout 120 python profiles/rows_timing.py; done
for r in 37 56 74; do echo rows_intra $r; MPTC_WAVE_ROWS_INTRA=$r timeout 200 python profiles/micro/k2_ab.py | cut -c1-150; done
