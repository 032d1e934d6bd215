#!/bin/bash
mkdir -p gpurun_out
for spec in "" "tile_steps=40" "min_blocks=1" "min_blocks=3" "block_threads=128 min_blocks=3" "block_threads=128 min_blocks=4" "block_threads=128 min_blocks=2" "block_threads=384 min_blocks=1" "block_threads=512 min_blocks=1" "tile_steps=40 min_blocks=1" "tile_steps=40 block_threads=128 min_blocks=3"; do
  echo "== $spec"
  python tools/run_cfg.py c3 5 $spec 2>&1 | tail -1
done
python tools/run_cfg.py c3t 5 2>&1 | tail -1
