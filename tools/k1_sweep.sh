#!/bin/bash
# Development probe: K1 staging / residency variants (UAVB_K1_VARIANT) on 10^6 config-2 missions.
for v in 0 1 2 3 4 5 6 7; do echo -n "variant $v: "; UAVB_K1_VARIANT=$v python tools/k1_probe.py | tail -1; done
