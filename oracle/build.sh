#!/bin/sh
# Build the CPU oracle (test infrastructure only).  -ffp-contract=off: the reference never fuses a*b+c.
set -e
cd "$(dirname "$0")"
mkdir -p _build
gcc -O2 -fno-fast-math -ffp-contract=off -fopenmp -shared -fPIC -o _build/librt_oracle.so rt_oracle.c -lm
echo "built oracle/_build/librt_oracle.so"
