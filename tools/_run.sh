(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6)
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" | tail -1
