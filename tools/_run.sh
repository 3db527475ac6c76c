(timeout 900 python -m pytest tests -m gpu -x -q -k "sharded_columns" 2>&1 | tail -6)
