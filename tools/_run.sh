(timeout 1200 python -m pytest tests -m gpu -x -q -k "non_power or arbitrary" 2>&1 | tail -15)
