#!/bin/bash
# Round artefacts in one GPU call: bench lines (not under a profiler), then the ncu launch lists of the same programs,
# then one `ncu --set full` capture of the conv_tc / wgrad_tc launches of one eager pass.  Usage: capture_profiles.sh <tag>
set -u
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
timeout 900 python bench.py > $out/${tag}_bench_infer.json 2> $out/${tag}_bench_infer.err
timeout 600 python bench.py --workload train --no-extra > $out/${tag}_bench_train.json 2> $out/${tag}_bench_train.err
timeout 600 python bench.py --workload infer1024 > $out/${tag}_bench_infer1024.json 2> $out/${tag}_bench_infer1024.err
timeout 300 python scripts/prof_infer.py > $out/${tag}_plain_infer.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_infer_launches.csv python scripts/prof_infer.py > $out/${tag}_ncu_infer.log 2>&1
timeout 300 python scripts/prof_train.py > $out/${tag}_plain_train.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_train_launches.csv python scripts/prof_train.py > $out/${tag}_ncu_train.log 2>&1
# full counters: the 23 conv_tc launches of the second inference pass, six wgrad_tc launches of the second training step.
# The .ncu-rep files stay on the box (gpurun_out is capped at 64 MiB): only their raw pages travel back as CSV.
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__block_size,launch__grid_size,launch__shared_mem_per_block_dynamic,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum
timeout 900 ncu --set full --clock-control none -k regex:conv_tc -s 23 -c 23 -o /tmp/${tag}_conv_tc_infer python scripts/prof_infer.py > $out/${tag}_ncu_full_infer.log 2>&1
ncu -i /tmp/${tag}_conv_tc_infer.ncu-rep --page raw --csv --metrics $M > $out/${tag}_conv_tc_infer_ncu.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none -k regex:wgrad_tc -s 20 -c 6 -o /tmp/${tag}_wgrad_tc_train python scripts/prof_train.py > $out/${tag}_ncu_full_train.log 2>&1
ncu -i /tmp/${tag}_wgrad_tc_train.ncu-rep --page raw --csv --metrics $M > $out/${tag}_wgrad_tc_train_ncu.csv 2>/dev/null
ls -la $out | grep ${tag}
