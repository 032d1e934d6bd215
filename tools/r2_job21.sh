#!/bin/bash
python tools/run_cfg.py c2cp 10 | tail -1
python -m pytest tests -m gpu -q -x -k "cp or shift or golden or examples or rare" 2>&1 | tail -3
