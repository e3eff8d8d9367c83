#!/bin/bash
# GPU box: `ncu --set full` capture of the three kernels of one utterance-CMVN step on the config-2
# batch with a WARM L2 (--cache-control none: the apply kernel is designed to hit in L2).
#   tools/ncu_step.sh NAME   ->  gpurun_out/NAME.ncu-rep (fbank, finalize, apply of one step)
name=$1
ncu --set full --clock-control none --cache-control none --import-source on \
    -k regex:'fbank_tile_kernel|finalize_utt_kernel|apply_kernel' -s 180 -c 3 \
    -o gpurun_out/$name -f python tools/quick_time.py $name > gpurun_out/$name.log 2>&1
tail -2 gpurun_out/$name.log
