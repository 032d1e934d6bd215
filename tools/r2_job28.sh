#!/bin/bash
python tools/run_cfg.py c3 5 | tail -1
python tools/run_cfg.py c3 5 ntp_direct=2 | tail -1
python tools/run_cfg.py c3 5 ntp_direct=4 | tail -1
python -m pytest tests/test_gpu_paths.py -m gpu -q -x -k "heston or jump or identical or ragged" 2>&1 | tail -3
