#!/bin/bash
# round 2, last session: deformable kernel back at the committed version (the 32-bit tap stride / hoisted image pointer
# variant measured 8-10 % slower): confirm
timeout 300 python scripts/kernel_bench.py deform 2>&1 | tail -5
