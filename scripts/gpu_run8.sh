python -m pytest tests/test_bands.py -m gpu -x -q 2>&1 | tail -3
