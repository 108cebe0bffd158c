#!/bin/bash
# timing experiments on the stem kernels (LEOD_STEM_DEBUG bits, see kernels_stem.cu)
for d in "$@"; do LEOD_STEM_DEBUG=$d python tools/stem_bench.py 2>&1 | tail -1; done
