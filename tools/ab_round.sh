timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
bash tools/ab.sh r4 m3 m5 f8 f7 fused 2>&1
B200JPEG_LIB_VARIANT=fused timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_transcode.py tests/test_raw_data.py -m gpu -x -q 2>&1 | tail -2
