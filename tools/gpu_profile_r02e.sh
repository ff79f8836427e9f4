TAG=r02e
O=gpurun_out
NCU="ncu --clock-control none"
COOP="python bench.py --steps 2 --warmup 3 --no-overlap --no-extras --no-cpu-baseline"
VPT="python bench.py --workload vpt --classes 102 --steps 2 --warmup 3 --no-overlap --no-extras --no-cpu-baseline"
$NCU --metrics gpu__time_duration.sum --launch-skip 700 -c 900 --csv --log-file $O/${TAG}_launches.csv $COOP > $O/${TAG}_ncu_coop.log 2>&1
$NCU --metrics gpu__time_duration.sum --launch-skip 700 -c 900 --csv --log-file $O/${TAG}_vpt_launches.csv $VPT > $O/${TAG}_ncu_vpt.log 2>&1
$NCU --set full --import-source on -k regex:attn_bwd_tc_kernel --launch-skip 30 -c 1 -f -o $O/${TAG}_attn_bwd_tc $VPT > $O/${TAG}_cap_attn_bwd_tc.log 2>&1
$NCU --set full --import-source on -k regex:attn_fwd_tc_kernel --launch-skip 30 -c 1 -f -o $O/${TAG}_attn_fwd_tc_l66 $VPT > $O/${TAG}_cap_attn_fwd_tc_l66.log 2>&1
ls -la $O/${TAG}_*.ncu-rep | awk '{print $5, $9}'
