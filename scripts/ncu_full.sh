#!/usr/bin/env bash
# Round-2 ncu evidence (one GPU, ONE gpurun call):
#   1. launch list of the bench command (time + DRAM bytes per launch)  -> gpurun_out/r02_launches.csv
#   2. --set full captures of the dominant kernels                        -> gpurun_out/r02_full_<name>.ncu-rep
# In one eager step the launches of a kernel are ordered as the net executes, so `-s` picks a layer:
#   k_conv_halo: forward launches 0..21 (up_7 = #18, up_9 = #20, deeper_10 = #10), then dgrad (up_9 = #23);  elementwise backward kernels start with the 256^2 layers.
set -u
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --profile > gpurun_out/r02_plain.log 2>&1 || { echo "plain bench run failed"; exit 1; }
# 2 eager steps + (1 + 5) replays of warm-up = 8 x 173 launches before the two timed replays
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1300 -c 450 --csv \
    --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --profile > gpurun_out/r02_ncu_list.log 2>&1
echo "launch list rc=$?"
run() {  # name script-args regex skip count
  ncu --set full --clock-control none --import-source on -k "regex:$3" -s "$4" -c "$5" -f -o gpurun_out/r02_full_$1 \
      python scripts/one_step.py $2 > gpurun_out/r02_ncu_full_$1.log 2>&1
  echo "$1 rc=$?"
}
python scripts/one_step.py den > gpurun_out/r02_plain_one_step.log 2>&1 || { echo "plain one_step failed"; exit 1; }
run conv_fwd_up9   den 'k_conv_halo'      20 1
run conv_dgrad_up9 den 'k_conv_halo'      23 1
run conv_fwd_up7   den 'k_conv_halo'      18 1
run conv_fwd_d10   den 'k_conv_halo'      10 1
run wgrad_up9      den 'k_wgrad_alias'     2 1
run pad_act_bwd    den 'k_pad_act_bwd'     2 1
run bn_bwd_apply   den 'k_bn_bwd_apply'    2 1
run bn_act_pad_fwd den 'k_bn_act_pad_fwd' 16 1
run cat_up_fwd     den 'k_cat_up_fwd'      4 1
run kl_reparam     den 'k_kl_reparam'      0 1
python scripts/one_step.py ct > gpurun_out/r02_plain_one_step_ct.log 2>&1 && {
run radon_fwd      ct  'k_radon_fwd'       0 1
run radon_bwd      ct  'k_radon_bwd'       0 1
}
