mkdir -p gpurun_out
B200JPEG_LIB_VARIANT=split timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/t7.log 2>&1; echo "pytest(split) rc=$?"; tail -n 3 gpurun_out/t7.log
tools/ab2.sh main split main split 2>&1 | tee gpurun_out/ab_r03_split.txt
