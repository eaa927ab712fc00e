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
# full counters: the 23 conv_tc launches of the second inference pass, the wgrad_tc launches of the second training step
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 23 -c 23 -o $out/${tag}_conv_tc_infer python scripts/prof_infer.py > $out/${tag}_ncu_full_infer.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc -s 20 -c 20 -o $out/${tag}_wgrad_tc_train python scripts/prof_train.py > $out/${tag}_ncu_full_train.log 2>&1
ls -la $out | grep ${tag}
