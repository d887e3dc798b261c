#!/bin/bash
# Round-1 profiling pass (run on the GPU box through gpurun, one GPU):
#   1. bench.py with one GOP lane  -> per-stage device times without overlap between lanes
#   2. ncu launch list of the bench command (gpu__time_duration per launch)
#   3. ncu --set full of the search kernels on the short 1080p workload (profiles/prof_cmd.py)
# Outputs land in gpurun_out/; the condensed summaries are committed under profiles/.
TAG=${1:-r1_v3}
mkdir -p gpurun_out
MPTC_LANES=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_lanes1.json 2> gpurun_out/${TAG}_bench_lanes1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launches.log 2>&1
MPTC_LANES=1 ncu --set full --clock-control none --import-source on \
    -k regex:'k_inter_search_tiled|k_intra_wavefront_tiled|k_intra_sparse' --launch-skip 4 -c 4 -f -o gpurun_out/${TAG}_search \
    python profiles/prof_cmd.py > gpurun_out/${TAG}_search.log 2>&1
MPTC_LANES=1 ncu --set full --clock-control none --import-source on \
    -k regex:'k_dxt1_fit|k_endpoint_planes|k_compact_unique' --launch-skip 3 -c 3 -f -o gpurun_out/${TAG}_stream \
    python profiles/prof_cmd.py > gpurun_out/${TAG}_stream.log 2>&1
tail -3 gpurun_out/${TAG}_bench_lanes1.json gpurun_out/${TAG}_search.log gpurun_out/${TAG}_stream.log
