#!/bin/bash
# Oracle helper (TEST INFRASTRUCTURE ONLY): build a scratch run directory for the reference binary.
# The reference writes its outputs next to the deck and opens opacity files relative to the CWD
# (io.cpp:122,145), and /root/reference is read-only, so decks are copied and inputdata/ is linked.
# usage: make_rundir.sh <rundir> [<reference root>]
set -e
RUN="$1"; REF="${2:-/root/reference}"
mkdir -p "$RUN"
for d in runs_mdot_vs_euv_plot runs_searth test_files test2 c2ray runs_earth PBC_model runs_convect runs_hjescape runs_schulik2020 friction_debug; do
    [ -d "$REF/$d" ] && { rm -rf "$RUN/$d"; cp -r "$REF/$d" "$RUN/$d"; }
done
cp "$REF/test_files/irradiation.spc" "$RUN/test2/"   # SEarth_opacities_1s22b.par names no species file
ln -sfn "$REF/inputdata" "$RUN/inputdata"
echo "$RUN ready"
