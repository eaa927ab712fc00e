#!/bin/bash
# Round artefacts in one GPU call: bench lines (not under a profiler), then the ncu launch lists of the same programs, then
# `ncu --set full` captures of the conv_tc launches of one inference pass and of one training step (DRAM traffic for
# roofline.traffic, tensor-pipe / operand-fetch / issue counters per launch).  Usage: capture_profiles.sh <tag>
set -u
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
timeout 1200 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
timeout 600 python bench.py --workload train --no-extra > $out/${tag}_bench_train.json 2> $out/${tag}_bench_train.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_reference_arm.err
timeout 300 python scripts/prof_infer.py > $out/${tag}_plain_infer.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_infer_launches.csv python scripts/prof_infer.py > $out/${tag}_ncu_infer.log 2>&1
timeout 300 python scripts/prof_train.py > $out/${tag}_plain_train.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_train_launches.csv python scripts/prof_train.py > $out/${tag}_ncu_train.log 2>&1
# full counters.  Inference: 4 passes x 23 conv_tc launches (constructor warm-ups + 2), the last pass = -s 69 -c 23.
# Training (UNPP_WGRAD_STREAM=0): the constructor's warm-up step + 2 steps of 57 conv_tc / 26 wgrad_tc launches; the last step.
# The .ncu-rep files stay on the box (gpurun_out is capped at 64 MiB): only selected columns of their raw pages travel back.
timeout 900 ncu --set full --clock-control none -k regex:conv_tc -s 69 -c 23 -o /tmp/${tag}_conv_tc_infer python scripts/prof_infer.py > $out/${tag}_ncu_full_infer.log 2>&1
ncu -i /tmp/${tag}_conv_tc_infer.ncu-rep --page raw --csv > /tmp/${tag}_conv_tc_infer_raw.csv 2>/dev/null
python scripts/ncu_select.py /tmp/${tag}_conv_tc_infer_raw.csv $out/${tag}_conv_tc_infer_ncu.csv
UNPP_WGRAD_STREAM=0 timeout 1500 ncu --set full --clock-control none -k regex:conv_tc -s 114 -c 57 -o /tmp/${tag}_conv_tc_train python scripts/prof_train.py > $out/${tag}_ncu_full_train.log 2>&1
ncu -i /tmp/${tag}_conv_tc_train.ncu-rep --page raw --csv > /tmp/${tag}_conv_tc_train_raw.csv 2>/dev/null
python scripts/ncu_select.py /tmp/${tag}_conv_tc_train_raw.csv $out/${tag}_conv_tc_train_ncu.csv
UNPP_WGRAD_STREAM=0 timeout 900 ncu --set full --clock-control none -k regex:wgrad_tc -s 52 -c 26 -o /tmp/${tag}_wgrad_tc_train python scripts/prof_train.py > $out/${tag}_ncu_full_wgrad.log 2>&1
ncu -i /tmp/${tag}_wgrad_tc_train.ncu-rep --page raw --csv > /tmp/${tag}_wgrad_tc_train_raw.csv 2>/dev/null
python scripts/ncu_select.py /tmp/${tag}_wgrad_tc_train_raw.csv $out/${tag}_wgrad_tc_train_ncu.csv
ls -la $out | grep ${tag}
