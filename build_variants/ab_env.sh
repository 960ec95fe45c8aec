#!/bin/bash
# build_variants/ab_env.sh LIB "VAR=a" "VAR=b" ...: time the iiwa step kernel of one library under different environments
LIB=$1; shift
for rep in 1 2; do
  for e in "$@"; do
    echo -n "$e: "; env $e ATACOM_B200_LIB=$PWD/build_variants/$LIB.so timeout 120 python profiles/profile_step.py time 2>&1 | grep "us/launch\|Error\|error" | cut -c1-90
  done
done
