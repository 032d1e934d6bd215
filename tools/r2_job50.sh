#!/bin/bash
for i in 1 2; do
SDE_B200_DEFINES="SDE_RES_GAMMA_SPECIALISE=0" python tools/run_cfg.py c2 20 | tail -1
python tools/run_cfg.py c2 20 | tail -1
done
SDE_B200_DEFINES="SDE_RES_GAMMA_SPECIALISE=0" python tools/energy_variants.py c2 150
python tools/energy_variants.py c2 150
SDE_B200_DEFINES="SDE_RES_GAMMA_SPECIALISE=0" python tools/energy_variants.py c2 150
python tools/energy_variants.py c2 150
python -m pytest tests/test_gpu_paths.py tests/test_golden.py tests/test_gpu_examples.py -m gpu -q -x 2>&1 | tail -3
