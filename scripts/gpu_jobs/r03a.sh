#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/variants.py t2b_20480 t2b_24576 t2b_28672 > gpurun_out/r03a_variants.log 2>&1
echo "variants rc=$?"; grep -E "variant|MISMATCH|exited|rror" gpurun_out/r03a_variants.log | cut -c1-330
