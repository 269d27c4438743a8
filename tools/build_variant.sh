#!/bin/bash
# Experiment builds: tools/build_variant.sh NAME "EXTRA nvcc flags"  →  build/variants/NAME.so  (select with MB_LIB=build/variants/NAME.so)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; EXTRA=$2
W=/tmp/mbvar_$NAME
rm -rf $W; mkdir -p $W/muscade.jl_b200 $ROOT/build/variants
cp -r $ROOT/muscade.jl_b200/csrc $W/muscade.jl_b200/csrc
cp -r $ROOT/include $W/include
make -C $W/muscade.jl_b200/csrc clean -s
make -C $W/muscade.jl_b200/csrc -j8 -s EXTRA="$EXTRA"
cp $W/muscade.jl_b200/libmuscade_b200.so $ROOT/build/variants/$NAME.so
grep -h -A1 "beam_static_sym" $W/muscade.jl_b200/csrc/beam_inst_10.ptxas.log | grep -i "spill\|Used" | head -3
