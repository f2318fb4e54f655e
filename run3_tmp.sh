python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -12
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 5 --warmup 2 --no-cpu 2>&1 | tail -c 600
