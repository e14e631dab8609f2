#!/bin/bash
# Builds every reference (Kokkos-OpenMP CPU) binary the parity tests and bench.py use.
# Needs /root/reference; the outputs under oracle/_ref/ travel to the GPU box via gpurun.
# `build_all.sh core` builds only what bench.py and smoke() use; the rest are the generators of tests/golden/short.
set -e
cd "$(dirname "$0")"
J=${J:-8}
b() { [ -x ../_ref/$1/idefix ] || make -s -j$J NAME=$1 SETUP=$2 MHD=$3 ORDER=$4 LOOP=$5 VP=$6; }
make -s -j$J kokkos
# core set: the CPU baseline of bench.py (ot3d_plm) and the smoke / headline parity cases
b ot3d_plm     test/MHD/OrszagTang3D      YES 2
b sod_plm      test/HD/sod                NO  2
b ot2d_plm     test/MHD/OrszagTang        YES 2 TPX
if [ "$1" = core ]; then echo "done (core set; run without argument for every fixture generator)"; exit 0; fi
b lw_ppm       test/MHD/LinearWaveTest    YES 4
b lw_plm       test/MHD/LinearWaveTest    YES 2
b lw_limo3     test/MHD/LinearWaveTest    YES 3
b ot2d_limo3   test/MHD/OrszagTang        YES 3 TPX
b ot2d_ppm     test/MHD/OrszagTang        YES 4 TPX
b mhdsod_plm   test/MHD/sod               YES 2
# EVOLVE_VECTOR_POTENTIAL builds (the reference's own OrszagTang set-ups carry the #ifdef branches)
b ot2d_plm_vp  test/MHD/OrszagTang        YES 2 TPX 1
b ot3d_plm_vp  test/MHD/OrszagTang3D      YES 2 "" 1
b sod_limo3    test/HD/sod                NO  3
b sod_ppm      test/HD/sod                NO  4
P=$(cd ../../idefix_b200/host/problems && pwd)
b sph_limo3    $P/SphericalMHD            YES 3
b sph_plm      $P/SphericalMHD            YES 2
b axisfluxtube_plm test/MHD/AxisFluxTube   YES 2
b axis_plm     $P/AxisMHD                 YES 2
b axis25d_plm  $P/Axis25D                 YES 2 TPX
b axis_limo3   $P/AxisMHD                 YES 3
b axis_plm_vp  $P/AxisMHD                 YES 2 "" 1
b polar_plm    $P/PolarMHD                YES 2
b sodiso_plm   $P/sodIso                  NO  2
b sodstrong_plm $P/sodStrong              NO  2
b othooks_plm  $P/otHooks                 YES 2 TPX
b iso25d_plm   $P/IsoMHD25D               YES 2 TPX
b iso25d_limo3 $P/IsoMHD25D               YES 3 TPX
b sph25d_plm   $P/Spherical25D            YES 2 TPX
b polarhd_ppm  $P/PolarHD2D               NO  4 TPX
b polarhd_limo3 $P/PolarHD2D              NO  3 TPX
echo done
