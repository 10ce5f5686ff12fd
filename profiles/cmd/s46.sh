python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dropblock_masks" -s 2>&1 | tail -12
