#!/bin/bash
for c in c5 c5p; do
  SDE_B200_NSTAGE=2 python tools/run_cfg.py $c 3 | tail -1
  python tools/run_cfg.py $c 3 | tail -1
  SDE_B200_NSTAGE=2 python tools/run_cfg.py $c 3 | tail -1
  python tools/run_cfg.py $c 3 | tail -1
done
