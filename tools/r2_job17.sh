#!/bin/bash
python tools/run_cfg.py c2 20 | tail -1
SDE_B200_DEBUG_NOSTORE=1 python tools/run_cfg.py c2 20 | tail -1
SDE_B200_DEBUG_NOCOMPUTE=1 python tools/run_cfg.py c2 20 | tail -1
python tools/run_cfg.py c2 20 | tail -1
