cd /root/repo
python tools/ncu_target.py > gpurun_out/ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"qconv_umma|cast_bf16" -c 16 -o gpurun_out/prof_conv_r1 -f python tools/ncu_target.py > gpurun_out/ncu_run.log 2>&1
echo "ncu rc=$?"; tail -5 gpurun_out/ncu_run.log; ls -la gpurun_out/*.ncu-rep
