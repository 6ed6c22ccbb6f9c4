cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1; echo "pytest all rc=$? $(tail -1 gpurun_out/pytest_all.log)"; grep -E "^FAILED|^E  " gpurun_out/pytest_all.log | cut -c1-300 | head -20
timeout 600 python tools/kprof.py --batch 1 > gpurun_out/kprof_b1.log 2>&1; echo "kprof rc=$?"; cat gpurun_out/kprof_b1.log | cut -c1-200
timeout 600 python tools/kprof.py --batch 4 --layers cnn1,tcn3,tcn1 > gpurun_out/kprof_b4.log 2>&1; echo "kprof rc=$?"; cat gpurun_out/kprof_b4.log | cut -c1-200
