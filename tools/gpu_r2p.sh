#!/bin/bash
# Round 2, GPU call P: where do the two-axes 16^3 cases differ from the reference dumps? (tools/axis_diag.py)
mkdir -p gpurun_out
timeout 300 python tools/axis_diag.py axis_hlld_contact_plm axis_hlld_ucthlld_plm axis_hll_contact_plm_rk3 axis_south_hlld_contact_plm axis_hlld_contact_plm_vp > gpurun_out/r2p.log 2>&1
grep -v "equal$" gpurun_out/r2p.log | cut -c1-330 | tail -90
