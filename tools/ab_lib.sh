#!/bin/bash
# same-box A/B of two builds of the library: tools/ab_lib.sh old.so new.so -> eval forward ms for each, twice
for rep in 1 2; do
  for lib in "$@"; do
    echo -n "$(basename $lib): "; MIPHEI_B200_LIB=$PWD/$lib python tools/time_eval.py 16 2>&1 | tail -1
  done
done
