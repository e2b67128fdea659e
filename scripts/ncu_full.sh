#!/usr/bin/env bash
# ncu --set full captures of the dominant kernels (one GPU).  Reports land in gpurun_out/r01_full_<name>.ncu-rep.
# In one eager step the launches of a kernel are ordered as the net executes, so `-s` picks a layer:
#   k_conv_halo: forward launches 0..25 (up_9 = #23), then dgrad;  elementwise backward kernels start with the 256^2 layers.
set -u
run() {  # name regex skip count
  ncu --set full --clock-control none --import-source on -k "regex:$2" -s "$3" -c "$4" -f -o gpurun_out/r01_full_$1 \
      python scripts/one_step.py > gpurun_out/ncu_full_$1.log 2>&1
  echo "$1 rc=$?"
}
run conv_fwd_up9   'k_conv_halo'      23 1
run conv_dgrad_up9 'k_conv_halo'      28 1
run conv_fwd_up7   'k_conv_halo'      21 1
run wgrad_up9      'k_wgrad_alias'     2 1
run pad_act_bwd    'k_pad_act_bwd'     2 1
run bn_bwd_apply   'k_bn_bwd_apply'    2 1
run bn_act_pad_fwd 'k_bn_act_pad_fwd' 16 1
run cat_up_fwd     'k_cat_up_fwd'      4 1
run kl_reparam     'k_kl_reparam'      0 1
run sample_weights 'k_sample_weights'  0 1
run adamw          'k_adamw'           0 1
