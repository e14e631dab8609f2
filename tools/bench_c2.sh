#!/bin/bash
# BASELINE config 2 through the C++ host: 2-D Orszag-Tang 1024^2, HLLD + PLM + UCT-HLLD, RK2 (`./idefix -nowrite`).
# Usage: tools/bench_c2.sh [cycles] [n]
set -e
cyc=${1:-200}; n=${2:-1024}
root=$(cd "$(dirname "$0")/.." && pwd)
work=$(mktemp -d)
make -s -C $root/idefix_b200/host PROBLEM=$root/idefix_b200/host/problems/OrszagTang MHD=YES ORDER=2 OUT=$work/idefix
sed -e "s/^X1-grid.*/X1-grid 1 0.0 $n u 1.0/" -e "s/^X2-grid.*/X2-grid 1 0.0 $n u 1.0/" -e "s/^log.*/log 100/" \
    $root/idefix_b200/host/problems/OrszagTang/idefix.ini > $work/idefix.ini
cd $work && ./idefix -i idefix.ini -nowrite -maxcycles $cyc | grep -E "Main: (Perfs|Reached|Completed)"
rm -rf $work
