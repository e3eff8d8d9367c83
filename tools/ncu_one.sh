#!/bin/bash
# GPU box: one `ncu --set full` capture of the fbank kernel on the config-2 batch.
#   tools/ncu_one.sh NAME [JS2T_LIB]   ->  gpurun_out/NAME.ncu-rep
name=$1
[ -n "$2" ] && export JS2T_LIB=$2
ncu --set full --clock-control none --import-source on -k regex:fbank_tile_kernel -s 3 -c 1 \
    -o gpurun_out/$name -f python tools/quick_time.py $name > gpurun_out/$name.log 2>&1
tail -2 gpurun_out/$name.log
