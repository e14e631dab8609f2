#!/bin/bash
# BASELINE config 5 through the C++ host: spherical 3-D MHD (log-r grid, geometric source terms), Roe + LimO3 +
# uct_contact, RK2, 384 x 192 x 384 per GPU, `./idefix -nowrite -maxcycles N` exactly as a reference user would
# time it ("Main: Perfs are ... cell updates/second", main.cpp:237-274).  Usage: tools/bench_c5.sh [cycles] [n1 n2 n3]
set -e
cyc=${1:-20}; n1=${2:-384}; n2=${3:-192}; n3=${4:-384}
root=$(cd "$(dirname "$0")/.." && pwd)
work=$(mktemp -d)
make -s -C $root/idefix_b200/host PROBLEM=$root/idefix_b200/host/problems/SphericalMHD MHD=YES ORDER=3 OUT=$work/idefix
sed -e "s/^X1-grid.*/X1-grid 1 1.0 $n1 l 4.0/" -e "s/^X2-grid.*/X2-grid 1 0.785398163397448 $n2 u 2.356194490192345/" \
    -e "s/^X3-grid.*/X3-grid 1 0.0 $n3 u 6.283185307179586/" -e "s/^log.*/log 10/" \
    $root/idefix_b200/host/problems/SphericalMHD/idefix.ini > $work/idefix.ini
cd $work && ./idefix -i idefix.ini -nowrite -maxcycles $cyc | grep -E "TimeIntegrator: +[0-9]|Main: (Perfs|Reached|Completed)|Fluid:"
rm -rf $work
