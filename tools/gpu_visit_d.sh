#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for cfg in mid_rotcam mid; do
for tc in 1 0; do
  FDGS_TILE_CULL=$tc python tools/grad_noise_probe.py $cfg aux 2>&1 | grep -E "tile_cull|dL_dts|dL_dscales|dL_drot |dL_dcov3D|dL_dmeans3D"
  FDGS_TILE_CULL=$tc FDGS_BLEND_BWD_V1=1 python tools/grad_noise_probe.py $cfg aux 2>&1 | grep -E "tile_cull|dL_dts|dL_dscales|dL_drot |dL_dcov3D|dL_dmeans3D"
done; done > $OUT/r02_v23_noise_probe.txt 2>&1
cat $OUT/r02_v23_noise_probe.txt
