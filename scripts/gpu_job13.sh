#!/bin/bash
# stress the MIS / multi-tile paths repeatedly on every variant to flush out races
for W in 2 4 1; do
  for i in 1 2 3; do
    DFB_TC_WPQ=$W timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "mis or config" 2>&1 | tail -1 | sed "s/^/WPQ=$W run $i: /"
  done
done
