for N in 2 4 6; do for S in 2 3 4; do echo "near $N split $S"; MPTC_LIB=$PWD/profiles/debug/variants/lib_near$N.so MPTC_ROW_SPLIT=$S python profiles/rows_timing.py; done; done
