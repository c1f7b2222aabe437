python scripts/debug_bands.py 2>&1 | tail -40
echo ==== ldg
SWB_KERNEL=ldg python scripts/debug_bands.py 2>&1 | tail -40
