#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/variants.py t2mb4 t2cap3072 t2cap4096 > gpurun_out/r03i_variants.log 2>&1
echo "variants rc=$?"; grep -E "variant|MISMATCH|exited|rror" gpurun_out/r03i_variants.log | cut -c1-330
