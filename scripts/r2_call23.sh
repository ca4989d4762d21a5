#!/usr/bin/env bash
# round 2, call 29: sub-phase stamps of the prepare kernel's triangular inverse
set -x
mkdir -p gpurun_out
timeout 300 python scripts/gdn_scan_tune.py > gpurun_out/scan_tune14.log 2>&1; grep -A4 "v2 prepare" gpurun_out/scan_tune14.log | cut -c1-200
