timeout 300 python -m pytest tests/test_gpu_scan_context.py -x -q 2>&1 | tail -3
timeout 100 python tools/sc_one.py 100000 32 tile 20
timeout 100 python tools/sc_one.py 100000 256 tile 10
timeout 100 python tools/sc_one.py 20000 32 tile 20
timeout 100 python tools/sc_one.py 20000 32 stream 20
