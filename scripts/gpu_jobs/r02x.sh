#!/bin/bash
mkdir -p gpurun_out
PARITY=1 SIZES=1 timeout 900 python scripts/any_check.py 24 > gpurun_out/r02x_any_steal1.log 2>&1; grep -E "ANY-HIT|rays:|half|dragon any" gpurun_out/r02x_any_steal1.log
RAYS=incoherent PARITY=0 timeout 900 python scripts/any_check.py 24 > gpurun_out/r02x_any_inc_steal1.log 2>&1; grep -E "dragon any" gpurun_out/r02x_any_inc_steal1.log
