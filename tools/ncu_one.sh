#!/bin/bash
# One full ncu capture of a kernel (regex $1) while running "$2..." ; report -> gpurun_out/$3.ncu-rep
K=$1; OUT=$2; shift 2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s ${SKIP:-2} -c ${COUNT:-1} -f -o gpurun_out/$OUT "$@" > gpurun_out/$OUT.log 2>&1
tail -3 gpurun_out/$OUT.log
