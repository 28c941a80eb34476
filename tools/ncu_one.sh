#!/bin/bash
# usage: ncu_one.sh <kernel-regex> <skip> <count> <outname> -- <cmd...>
K=$1; S=$2; C=$3; O=$4; shift 5
mkdir -p gpurun_out
timeout -k 10 300 "$@" > gpurun_out/${O}_plain.log 2>&1 &&
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c $C -o gpurun_out/$O -f "$@" > gpurun_out/${O}_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/${O}_plain.log
