(timeout 1200 python -m pytest tests -m gpu -x -q -k "resident" 2>&1 | tail -25)
