timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "exchange" 2>&1 | grep -v "^$" | head -60
