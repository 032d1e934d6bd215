#!/bin/bash
python tools/energy_variants.py c2 150
SDE_B200_DEFINES='SDE_ST_HINT=".cs"' python tools/energy_variants.py c2 150
SDE_B200_DEFINES='SDE_ST_HINT=".wt"' python tools/energy_variants.py c2 150
SDE_B200_DEFINES='SDE_ST_HINT=".cg"' python tools/energy_variants.py c2 150
SDE_B200_DEFINES='SDE_ST_HINT=".L2::evict_first"' python tools/energy_variants.py c2 150 2>&1 | tail -1
python tools/energy_variants.py c2 150
