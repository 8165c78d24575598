set -x
python scripts/msdist_diag.py 2>&1 | grep -v "^variant" | tail -5
python -m pytest tests/test_gpu_production_samplers.py tests/test_gpu_parity.py -m gpu -q -s 2>&1 | grep -E "PARITY|passed|failed|FAILED|Error|KS" | cut -c1-600 | tail -40
