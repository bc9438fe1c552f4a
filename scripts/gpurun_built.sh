#!/bin/bash
# build-container helper: (re)build the library synchronously, refuse to go to the GPU with register spills in the tile
# kernels, then hand the command to gpurun.   usage: scripts/gpurun_built.sh <timeout-s> '<command>'
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()"
STACK=$(cuobjdump --dump-resource-usage unitaria_b200/libunitaria_b200.so 2>/dev/null | grep -A1 "k_tile_swz" | grep -o "STACK:[0-9]*" | cut -d: -f2 | sort -n | tail -1)
echo "max stack of k_tile_swz variants: $STACK bytes"
if [ "$STACK" -gt 64 ]; then echo "tile kernels spill ($STACK bytes of stack): fix before spending GPU time"; exit 1; fi
exec /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
