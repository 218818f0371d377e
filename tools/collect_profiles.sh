#!/bin/bash
# Runs on the GPU box (gpurun): refreshes every single-GPU measurement that profiles/ cites (round-2 names).  Each ncu
# capture follows a plain run of the same command that exited 0.  Outputs land in gpurun_out/prof (merged back by gpurun);
# copy what should be judged into profiles/.  The multi-GPU lines come from
#   gpurun --gpus N -- python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bench.py --gpus N
set -u
O=gpurun_out/prof
mkdir -p $O
python bench.py --steps 20 --warmup 5 > $O/r02_bench_1gpu.json 2> $O/bench_1gpu.err || echo "bench failed"
python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference_arm.json 2> $O/bench_reference_arm.err || echo "reference arm failed"
# launch list of the library's kernels (torch's data-generation kernels filtered out by name)
python bench.py --steps 2 --warmup 3 --e2e-steps 1 --configs-steps 1 > $O/bench_short.json 2> $O/bench_short.err && \
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"nq_scan|^k_" -c 600 --csv --log-file $O/r02_launches_bench_1gpu.csv \
      python bench.py --steps 2 --warmup 3 --e2e-steps 1 --configs-steps 1 > $O/ncu_launches.log 2>&1
# config 5 (the headline kernel), 200 M rows
FS_SCALE=0.2 python tools/full_size.py config5 > $O/full_size_config5.jsonl 2> $O/full_size.err && \
  FS_SCALE=0.2 ncu --set full --clock-control none --import-source on -k regex:nq_scan -s 1 -c 1 -f -o $O/r02_nq_scan_config5 \
      python tools/full_size.py config5 > $O/ncu_config5.log 2>&1
ncu -i $O/r02_nq_scan_config5.ncu-rep --page raw --csv > $O/r02_nq_scan_config5_ncu_full_raw.csv 2>/dev/null
# config 4: partition kernel (entry nq_scan), k_part_aggregate, k_finalize_groups at 200 M rows
python tools/full_size.py config4 >> $O/full_size_config5.jsonl 2>> $O/full_size.err && \
  ncu --set full --clock-control none --import-source on -k regex:"nq_scan|k_part_aggregate|k_finalize_groups" -s 3 -c 3 -f -o $O/r02_config4 \
      python tools/full_size.py config4 > $O/ncu_config4.log 2>&1
ncu -i $O/r02_config4.ncu-rep --page raw --csv > $O/r02_config4_partition_aggregate_finalize_ncu_full_raw.csv 2>/dev/null
# the config-5 kernel one change at a time (1 B rows), and the hand-written prototype of the same kernel
ROWS=1000000000 EXTRA="$(cat tools/config5_ablation.json)" python tools/sweep_config5.py > $O/r02_config5_ablation.jsonl 2> $O/config5_sweep.err
if [ -x tools/_build/proto5 ]; then
  tools/_build/proto5 2>/dev/null | grep -v '"variant": "base' > $O/r02_proto5_ablation.jsonl
  HOT=1 ONLY="simple w1 aos" tools/_build/proto5 2>/dev/null | grep -v "base\|nomnreg" | sed 's/"variant": "/"variant": "hot keys preloaded: /' >> $O/r02_proto5_ablation.jsonl
  for m in 1 2; do [ -x tools/_build/proto5_m$m ] && ONLY="simple w1 aos" tools/_build/proto5_m$m 2>/dev/null | grep -v "base\|nomnreg\|nocheck\|7040" | sed "s/\"variant\": \"/\"variant\": \"-DPROTO_MODE=$m: /" >> $O/r02_proto5_ablation.jsonl; done
fi
rm -f $O/*.ncu-rep
ls -la $O
