#!/bin/bash
# Developer tool (GPU box): time the step with the in-tree library and every build_variants/*.so
# usage: scripts/sweep.sh "<time_step args>" ...
for args in "$@"; do
  echo "## $args"
  python scripts/time_step.py $args
  for lib in build_variants/*.so; do
    [ -f "$lib" ] && PTOY_B200_LIB=$PWD/$lib python scripts/time_step.py $args
  done
done
