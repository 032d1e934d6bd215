#!/bin/bash
python tools/run_cfg.py c3 5 | tail -1
SDE_B200_DEFINES="SDE_DEBUG_NOBARRIER=1" python tools/run_cfg.py c3 5 | tail -1
python tools/run_cfg.py c3t 5 | tail -1
SDE_B200_DEFINES="SDE_DEBUG_NOBARRIER=1" python tools/run_cfg.py c3t 5 | tail -1
