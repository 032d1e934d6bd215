#!/bin/bash
python tools/energy_variants.py c2 150
SDE_B200_DEBUG_NOSTORE=1 python tools/energy_variants.py c2 150
SDE_B200_DEBUG_NOCOMPUTE=1 python tools/energy_variants.py c2 150
python tools/energy_variants.py c2 150
