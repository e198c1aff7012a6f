#!/bin/bash
for m in 0 1 2 3 4 7; do echo "wait_mode $m"; DGX_MARCH_WAIT=$m python tools/march_variants.py 7 50 v2eor 2>&1 | tail -1; done
