set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python scripts/steady.py cur 40000000
