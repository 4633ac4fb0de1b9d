#!/bin/bash
# Evidence of a round, in one GPU call (run under gpurun from the repo root):
#   1. bench.py, both arms, without a profiler          -> gpurun_out/rNN_bench*.json
#   2. the ncu launch list of the same bench command     -> gpurun_out/rNN_launches_bench.csv
#   3. one `ncu --set full` capture of the dominant kernels at 1 GiB (decode, hash parse, exact parse)
# usage: tools/profile_round.sh r02
set -u
R=${1:-r02}
O=gpurun_out
mkdir -p $O
timeout 600 python bench.py > $O/${R}_bench.json 2> $O/${R}_bench.err || echo "bench failed"
timeout 600 python bench.py --impl reference > $O/${R}_bench_reference_arm.json 2>> $O/${R}_bench.err || echo "reference arm failed"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^k_|sb200' --csv \
    --log-file $O/${R}_launches_bench.csv python bench.py --no-cpu-baseline > $O/${R}_bench_under_ncu.json 2>> $O/${R}_bench.err
timeout 600 ncu --set full --import-source on --clock-control none -k 'regex:k_decode_seg|k_copy_literal_blocks' -c 2 \
    -o $O/${R}_full_decode python tools/prof_decode.py --mib 1024 > /dev/null 2>> $O/${R}_bench.err
timeout 600 ncu --set full --import-source on --clock-control none -k 'regex:k_parse_hash_global|k_emit' -c 2 \
    -o $O/${R}_full_parse python tools/prof_run.py --mib 1024 --kind mixed --mode 0 > /dev/null 2>> $O/${R}_bench.err
timeout 600 ncu --set full --import-source on --clock-control none -k 'regex:k_parse_exact_global' -c 2 \
    -o $O/${R}_full_parse_exact python tools/prof_run.py --mib 1024 --kind lowent_random --mode 1 > /dev/null 2>> $O/${R}_bench.err
ls -la $O | tail -12
tail -3 $O/${R}_bench.err
