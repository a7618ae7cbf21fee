#!/bin/bash
# One B200: the bench line, the reference arm, the ncu launch list of the same command and ONE full capture of the sweep kernel.
# usage (on the GPU box, from the repo root): bash scripts/round_profile.sh r02d
set -u
P=${1:-rXX}
O=gpurun_out
mkdir -p $O
python bench.py --steps 20 --warmup 5 > $O/${P}_bench_n1.json 2> $O/${P}_bench_n1.err || { tail -5 $O/${P}_bench_n1.err; exit 1; }
python bench.py --impl reference --steps 3 --warmup 1 > $O/${P}_bench_reference_cpu.json 2> $O/${P}_bench_reference_cpu.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${P}_launches_bench_steps20.csv \
    python bench.py --steps 20 --warmup 5 > $O/${P}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pipe -s 1 -c 1 -f -o $O/${P}_k_pipe \
    python bench.py --steps 20 --warmup 5 > $O/${P}_ncu_full.log 2>&1
[ "${SKIP_BIN:-0}" = 1 ] || ncu --set full --clock-control none --import-source on -k regex:k_bin -s 2 -c 1 -f -o $O/${P}_k_bin \
    python bench.py --steps 20 --warmup 5 > $O/${P}_ncu_full_bin.log 2>&1
tail -1 $O/${P}_bench_n1.json | cut -c1-400
