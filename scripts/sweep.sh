#!/bin/bash
# usage: scripts/sweep.sh [time_das args...] ; runs every lib variant in qups_b200/variants plus the default build
cd "$(dirname "$0")/.."
echo "== default"; timeout 300 python scripts/time_das.py "$@" 2>&1 | tail -1
for f in qups_b200/variants/lib_*.so; do
  echo "== $f"; QUPS_B200_LIB=$PWD/$f timeout 300 python scripts/time_das.py "$@" 2>&1 | tail -1
done
