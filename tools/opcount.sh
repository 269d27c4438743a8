#!/bin/sh
# Frozen algorithmic flop counts of the element kernels' mathematics: builds tools/opcount.cpp (the product's device math headers with a counting scalar,
# host g++) and writes profiles/falg.json.  Re-run after any change to muscade.jl_b200/csrc/{dual,sdual,beam_math}.cuh and update BASELINE.md §5.
set -e
cd "$(dirname "$0")/.."
/usr/bin/g++ -O1 -std=c++17 -o /tmp/mb_opcount tools/opcount.cpp
/tmp/mb_opcount > profiles/falg.json
cat profiles/falg.json
