set -u
O=gpurun_out; TAG=r2end
timeout 400 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
  -k regex:qconv_cl_fprop_kernel --csv --log-file $O/${TAG}_ncu_traffic.csv python tools/ncu_step.py > $O/${TAG}_ncu_traffic.log 2>&1; echo "ncu traffic rc=$?"
i=0
for spec in "qconv_cl_fprop_kernel:14" "qconv_cl_wgrad_kernel:6" "first_layer_bwd_kernel:1" "attn_kernel:3"; do
  k="${spec%%:*}"; c="${spec##*:}"; i=$((i+1))
  SELDQ_PDL=0 SELDQ_SIDE_WGRAD=0 timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"$k" -c $c -o $O/${TAG}_ncu_full_$i -f python tools/ncu_step.py > $O/${TAG}_ncu_full_$i.log 2>&1; echo "ncu full $k rc=$?"
done
ls -la $O/${TAG}_ncu_full_* | cut -c30-
