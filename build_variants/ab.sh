#!/bin/bash
# build_variants/ab.sh NAME...: time the iiwa step kernel of each variant library, interleaved, on this box
for rep in 1 2; do
  for v in "$@"; do
    echo -n "$v: "; ATACOM_B200_LIB=$PWD/build_variants/$v.so timeout 120 python profiles/profile_step.py time 2>&1 | grep "us/launch" | cut -c1-60
  done
done
