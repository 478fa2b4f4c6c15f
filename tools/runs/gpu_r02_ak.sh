#!/bin/bash
for b in 4 2; do for s in 64 128 256 512; do echo "seg $s"; DCTC_SMALL_SEG=$s timeout 60 python tools/time_small.py $b 64 5; DCTC_SMALL_SEG=$s timeout 60 python tools/time_small.py $b 16 10; done; done
