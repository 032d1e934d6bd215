#!/bin/bash
python tools/energy_variants.py c2 150
SDE_B200_NO_WIDE_TABLE=1 python tools/energy_variants.py c2 150
python tools/energy_variants.py c2 150
python tools/energy_variants.py c2 150 block_threads=512
python tools/energy_variants.py c2 150 block_threads=640
python tools/energy_variants.py c2 150 block_threads=896
python tools/energy_variants.py c2 150
