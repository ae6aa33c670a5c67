python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 3 2>/dev/null | tail -1 | cut -c1-200
