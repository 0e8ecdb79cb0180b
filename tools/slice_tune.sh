#!/bin/bash
# Development aid (GPU box): K2 headline launch against the two constants of the decreasing-slice rule (rollout_kernels.cu), read from the
# environment by a development build:  python uav-autonomous-control_b200/build.py --variant dev -DUAVB_DEV  &&  tools/slice_tune.sh
lib=uav-autonomous-control_b200/lib/libuavb.so
cp $lib /tmp/libuavb_default.so
cp build/variants/libuavb_dev.so $lib
for part in 3 4 5 6 8; do for mn in 100 200 300 400; do
  echo -n "part=$part min=$mn: "; UAVB_SLICE_PART=$part UAVB_SLICE_MIN=$mn python tools/k2_probe.py 100000 6 | tail -1
done; done
cp /tmp/libuavb_default.so $lib
