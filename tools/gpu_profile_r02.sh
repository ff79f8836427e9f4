#!/bin/bash
# Round-2 profiling pass (run under gpurun, one GPU): launch lists of the CoOp and VPT steps and one `ncu --set full`
# capture per kernel family that had none in round 1.  Raw files → gpurun_out/; tools/summarize_profiles.py → profiles/.
TAG=${1:-r02a}
O=gpurun_out
NCU="ncu --clock-control none"
COOP="python bench.py --steps 2 --warmup 3 --no-overlap --no-extras --no-cpu-baseline"
VPT="python bench.py --workload vpt --classes 102 --steps 2 --warmup 3 --no-overlap --no-extras --no-cpu-baseline"
$NCU --metrics gpu__time_duration.sum --launch-skip 700 -c 900 --csv --log-file $O/${TAG}_launches.csv $COOP > $O/${TAG}_ncu_coop.log 2>&1
$NCU --metrics gpu__time_duration.sum --launch-skip 700 -c 900 --csv --log-file $O/${TAG}_vpt_launches.csv $VPT > $O/${TAG}_ncu_vpt.log 2>&1
cap() {  # name regex skip cmd…
  local name=$1 re=$2 skip=$3; shift 3
  $NCU --set full --import-source on --kernel-name-base demangled -k regex:"$re" --launch-skip $skip -c 1 -f -o $O/${TAG}_$name "$@" > $O/${TAG}_cap_$name.log 2>&1
}
if [ "$2" = "gemm" ]; then
cap gemm_fc '2cta_kernel<.int.1, .int.4>' 30 $COOP
cap gemm_resid '2cta_kernel<.int.1, .int.2>' 60 $COOP
cap gemm_act2 '2cta_kernel<.int.1, .int.5>' 30 $VPT
cap gemm_dgrad '2cta_kernel<.int.1, .int.0>' 60 $VPT
ls -la $O/${TAG}_*.ncu-rep | awk '{print $5, $9}'
exit 0
fi
cap attn_fwd_tc 'attn_fwd_tc_kernel' 30 $COOP
cap gemm_fc '2cta_kernel<.int.1, .int.4>' 30 $COOP
cap gemm_resid '2cta_kernel<.int.1, .int.2>' 60 $COOP
cap replay_par 'lb_replay_par_kernel' 3 $COOP
cap im2col_u8 'im2col_patch32_u8_kernel' 3 $COOP
cap assemble 'vit_assemble_lnpre_kernel' 3 $COOP
cap attn_bwd 'attn_bwd_kernel' 30 $VPT
cap ln_bwd 'layernorm_bwd_kernel' 60 $VPT
cap gemm_act2 '2cta_kernel<.int.1, .int.5>' 30 $VPT
cap gemm_dgrad '2cta_kernel<.int.1, .int.0>' 60 $VPT
cap prefix_grad 'prefix_grad_kernel' 2 $VPT
ls -la $O/${TAG}_*.ncu-rep | awk '{print $5, $9}'
