#!/bin/bash
# float64 pipeline step time (pipe_probe.py, both allocation orders) per variant library
for v in "$@"; do
  for mode in bench ab; do
    echo -n "$v $mode: "
    NPHUSL_CUDA_LIB=$PWD/build/var/lib_$v.so python profiles/tools/pipe_probe.py $mode | grep -E "raw events=0|per-20" | sed -E 's/per-20-step ms\/step: ([0-9.]+) ([0-9.]+) ([0-9.]+).* ([0-9.]+) ([0-9.]+)$/drift \2 \3 ... \4 \5/' | tr '\n' ' '
    echo
  done
done
