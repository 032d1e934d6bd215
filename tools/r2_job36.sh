#!/bin/bash
python -m pytest tests/test_gpu_paths.py -m gpu -q -x -k "bulk_copy" 2>&1 | tail -5
python tools/run_cfg.py c3 5 | tail -1
python tools/run_cfg.py c3 5 ntp_direct=5 | tail -1
python tools/run_cfg.py c3 5 ntp_direct=5 tile_steps=24 | tail -1
python tools/run_cfg.py c3 5 ntp_direct=5 tile_steps=40 | tail -1
