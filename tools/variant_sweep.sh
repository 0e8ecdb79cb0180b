#!/bin/bash
# Development aid (GPU box): run a probe command once per library variant built by `build.py --variant <tag> ...`
#   tools/variant_sweep.sh "python tools/k2_probe.py 100000 4" r128 r160 r192
cmd="$1"; shift
lib=uav-autonomous-control_b200/lib/libuavb.so
cp $lib /tmp/libuavb_default.so
echo "== default"; $cmd
for tag in "$@"; do
  cp build/variants/libuavb_$tag.so $lib
  echo "== $tag"; $cmd
done
cp /tmp/libuavb_default.so $lib
