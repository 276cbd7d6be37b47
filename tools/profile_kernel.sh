#!/bin/bash
# usage: tools/profile_kernel.sh <kernel regex> <skip> <count> <out name> [batches] [scene]
# one `ncu --set full` capture of a kernel of the bench workload (never a bench number)
ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c $3 -o gpurun_out/$4 -f python tools/profile_run.py ${5:-1} ${6:-Atrium} > gpurun_out/$4.log 2>&1
tail -2 gpurun_out/$4.log
