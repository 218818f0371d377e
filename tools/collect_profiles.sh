#!/bin/bash
# Runs on the GPU box (gpurun): refreshes every single-GPU measurement that profiles/ cites.  Each ncu capture follows a
# plain run of the same command that exited 0.  Outputs land in gpurun_out/prof (merged back by gpurun).
set -u
O=gpurun_out/prof
mkdir -p $O
python bench.py > $O/bench_1gpu.json 2> $O/bench_1gpu.err || echo "bench failed"
python bench.py --impl reference > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err || echo "reference arm failed"
python bench.py --steps 2 --warmup 3 > $O/bench_short.json 2> $O/bench_short.err && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench_config2.csv \
      python bench.py --steps 2 --warmup 3 > $O/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k nq_scan -s 6 -c 1 -f -o $O/nq_scan_config2 \
      python bench.py --steps 2 --warmup 3 > $O/ncu_config2.log 2>&1
ncu -i $O/nq_scan_config2.ncu-rep --page raw --csv > $O/nq_scan_config2_ncu_full_raw.csv 2>/dev/null
python tools/full_size.py > $O/full_size_configs345.jsonl 2> $O/full_size.err || echo "full_size failed"
ROWS=10000000,60000000 python tools/scan_perf.py config2 config2_sel1 config3 > $O/scan_perf_shapes.txt 2>&1
# config 5 on its real key distribution (Zipf 1.1 over 100 k permuted ranks), 200 M rows: knob-by-knob sweep, then ncu
ROWS=200000000 python tools/sweep_config5.py > $O/config5_sweep.jsonl 2> $O/config5_sweep.err && \
  ROWS=200000000 ONLY="default (all" ncu --set full --clock-control none --import-source on -k nq_scan -s 3 -c 1 -f -o $O/nq_scan_config5 \
      python tools/sweep_config5.py > $O/ncu_config5.log 2>&1
ncu -i $O/nq_scan_config5.ncu-rep --page raw --csv > $O/nq_scan_config5_direct_ncu_full_raw.csv 2>/dev/null
# config 4 (1 M groups, COUNT + SUM DISTINCT): first pass of the sliced bitmap scan, 50 M rows
FS_SCALE=0.25 python tools/full_size.py config4 > /dev/null 2>&1 && \
  FS_SCALE=0.25 ncu --set full --clock-control none --import-source on -k nq_scan -s 2 -c 1 -f -o $O/nq_scan_config4 python tools/full_size.py config4 > $O/ncu_config4.log 2>&1
ncu -i $O/nq_scan_config4.ncu-rep --page raw --csv > $O/nq_scan_config4_direct_bitmap_ncu_full_raw.csv 2>/dev/null
rm -f $O/*.ncu-rep
ls -la $O
