#!/bin/bash
# Profiling pass (run on the GPU box through gpurun, one GPU):
#   1. bench.py with one GOP lane  -> per-stage device times without overlap between lanes
#   2. ncu launch list of the bench command (gpu__time_duration per launch)
#   3. ncu --set full of the search kernels on the short 1080p workload (profiles/prof_cmd.py)
# Outputs land in gpurun_out/; the condensed summaries are committed under profiles/.
TAG=${1:-r1_v4}
mkdir -p gpurun_out
MPTC_LANES=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra-legs > gpurun_out/${TAG}_bench_lanes1.json 2> gpurun_out/${TAG}_bench_lanes1.err
# launch list: one GOP lane, so that one launch covers frame k of every GOP exactly as in the attribution pass of
# bench.py (with 4 lanes ncu serialises the 4 concurrent wavefront launches and their share quadruples)
MPTC_LANES=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra-legs > gpurun_out/${TAG}_launches.log 2>&1
MPTC_LANES=1 ncu --set full --clock-control none --import-source on \
    -k regex:'k_inter_search_wide|k_inter_search_tiled|k_intra_rows|k_intra_sparse' --launch-skip 4 -c 4 -f -o gpurun_out/${TAG}_search \
    python profiles/prof_cmd.py > gpurun_out/${TAG}_search.log 2>&1
MPTC_LANES=1 ncu --set full --clock-control none --import-source on \
    -k regex:'k_dxt1_fit|k_endpoint_planes|k_compact_unique|k_compact_count' --launch-skip 4 -c 4 -f -o gpurun_out/${TAG}_stream \
    python profiles/prof_cmd.py > gpurun_out/${TAG}_stream.log 2>&1
for K in k_inverse_planes k_dxt1_to_rgb k_dec_links; do
  ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip 1 -c 1 -f -o gpurun_out/${TAG}_$K \
      python profiles/prof_cmd.py > gpurun_out/${TAG}_$K.log 2>&1
done
tail -3 gpurun_out/${TAG}_bench_lanes1.json gpurun_out/${TAG}_search.log gpurun_out/${TAG}_stream.log
