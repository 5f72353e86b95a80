#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/variants.py r02j t1_512x2 t1_1024x1 > gpurun_out/r02j_variants.log 2>&1
echo "variants rc=$?"; grep -E "variant|MISMATCH|exited" gpurun_out/r02j_variants.log | cut -c1-420
