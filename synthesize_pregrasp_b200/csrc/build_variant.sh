#!/bin/bash
# Builds an experiment variant of the library: only the rollout translation units are recompiled with the given defines.
#   ./build_variant.sh <tag> [box|hull|both] <nvcc defines...>      ->  ../variants/libpregrasp_b200_<tag>.so
set -e
cd "$(dirname "$0")"
tag=$1; which=$2; shift 2
mkdir -p ../variants vobj
NVCC=/usr/local/cuda/bin/nvcc
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v"
box=pg_rollout.o; hull=pg_rollout_hull.o   # note: the default box object is now built with -fmad=false as well (Makefile)
if [ "$which" = box ] || [ "$which" = both ]; then $NVCC $ARCH $FLAGS -fmad=false "$@" -c -o vobj/box_$tag.o pg_rollout.cu 2> vobj/box_$tag.log & box=vobj/box_$tag.o; fi
if [ "$which" = hull ] || [ "$which" = both ]; then $NVCC $ARCH $FLAGS -fmad=false "$@" -c -o vobj/hull_$tag.o pg_rollout_hull.cu 2> vobj/hull_$tag.log & hull=vobj/hull_$tag.o; fi
wait
$NVCC $ARCH -shared -o ../variants/libpregrasp_b200_$tag.so pg_api.o $box $hull pg_rollout_phase.o pg_rollout_phase_hull.o pg_cost.o pg_cost_tc.o pg_mppi.o pg_dynfeas.o
grep -h "Used" vobj/*_$tag.log | head -2
