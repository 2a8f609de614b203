mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_txsynth.py tests/test_zbmac.py tests/test_gpu_parity.py -m gpu -x -q -k "txsynth or gpu_capture or zbmac or mac_summary or exchange" 2>&1 | tail -25
