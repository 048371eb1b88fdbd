#!/bin/bash
# usage (GPU box): tools/ncu_one.sh <tag> <kernel regex> [bench args]  -> one --set full capture (with source) of the
# kernel's 2nd launch in a one-step bench run: gpurun_out/<tag>.ncu-rep
tag=$1; rx=$2; shift 2
CMD="python bench.py --steps 1 --warmup 1 --no-cpu $*"
timeout 120 $CMD > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${tag}_plain.log; exit 1; }
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s 1 -c 1 \
    -o gpurun_out/${tag} -f $CMD > gpurun_out/${tag}_ncu.log 2>&1
ls -la gpurun_out/${tag}*
