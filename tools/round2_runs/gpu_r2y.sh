#!/bin/bash
# round 2: clock64 phase breakdown of bev_band / bev_bin (debug-timing build made on the box; the release .so is restored by the snapshot being discarded)
mkdir -p gpurun_out
SFA_DEBUG_TIMING=1 python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
for ring in 8 32; do for lanes in 1 2; do
echo "== ring $ring internal lanes $lanes"
SFA_DEBUG_TIMING=1 SFA_BEV_TILED_RING=$ring SFA_BEV_INTERNAL_LANES=$lanes python tools/band_timing.py
done; done
