timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
bash tools/ab.sh ar a r ar4 old4 ar6 ilp arilp kz occ8 enc sp spenc fused sort hist fmask fall all all2 2>&1
B200JPEG_LIB_VARIANT=fall timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
B200JPEG_LIB_VARIANT=spenc timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
B200JPEG_LIB_VARIANT=arilp timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
B200JPEG_LIB_VARIANT=fused timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_transcode.py -m gpu -x -q 2>&1 | tail -2
B200JPEG_LIB_VARIANT=all timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
