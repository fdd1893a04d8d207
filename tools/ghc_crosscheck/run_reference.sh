#!/bin/sh
# Runs the REAL reference (flannelhead/blackstar, Haskell) on the cross-check inputs.  Needs `stack`
# (resolver lts-13.16, GHC 8.6.4 -- stack.yaml of the reference) and a checkout of the reference.
#
#   usage: tools/ghc_crosscheck/run_reference.sh /path/to/blackstar-checkout INPUTS_DIR OUT_DIR
#
# INPUTS_DIR is what make_inputs.py wrote (stars.ppm + scenes/).  OUT_DIR receives prev-<scene>.png, the
# files `blackstar --preview` writes (app/Main.hs:86,97-101).  Then:
#   python tools/ghc_crosscheck/compare.py OUT_DIR
set -eu
REF=$(cd "$1" && pwd); IN=$(cd "$2" && pwd); mkdir -p "$3"; OUT=$(cd "$3" && pwd)
cd "$REF"
stack build
# PPM catalogue -> k-d tree file (app/GenerateTree.hs:11-29)
stack exec generate-tree -- "$IN/stars.ppm" "$OUT/stars.kdt"
# every scene, preview mode, no overwrite prompts (app/Main.hs:20-41)
for s in "$IN"/scenes/*.yaml; do
    stack exec blackstar -- --preview --force --starmap "$OUT/stars.kdt" --output "$OUT" "$s"
done
ls -l "$OUT"/prev-*.png
