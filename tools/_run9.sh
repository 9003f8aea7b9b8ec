mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/t6.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/t6.log
tools/final_evidence.sh r03
