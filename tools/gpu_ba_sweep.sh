#!/bin/bash
# batched-affine rounds sweep (development aid): device-resident phase times for R = 0..3
for spec in g1:20 g1:22 g1:24 g2:20 g1:18 g1:16; do
  for ba in 0 1 2 3; do
    echo -n "BA=$ba "; BA=$ba python tools/gpu_sizes.py $spec 2>&1 | tail -1
  done
done
