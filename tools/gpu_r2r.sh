#!/bin/bash
# round 2, GPU call R: knob sweep on the final kernels (env switches only)
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" python tools/gpu_rankshare.py 2>&1 | head -2 | tr '\n' ' '; echo; }
run X=1
run WRT_DEEP_SPLIT=6
run WRT_DEEP_SPLIT=7
run WRT_TRACE_BLOCKS=8
run WRT_TRACE_BLOCKS=9
run WRT_TRACE_BLOCKS=12
run WRT_SIDE_BLOCKS=6
run WRT_SIDE_BLOCKS=4
run WRT_REFILL=12
run WRT_REFILL=20
run WRT_REFILL=24
run WRT_SHADE0_SEPARATE=0
run WRT_CHUNK_DIV=8
run WRT_CHUNK_DIV=32
