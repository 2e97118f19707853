timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python profiles/c5_fleet.py --radars 128 --handles 8 --only 4 --passes 5 2>&1 | python -c "import sys,json; d=json.load(sys.stdin); print([(r['window'], r['hop'], round(r['seconds_per_pass_max_over_ranks']*1e3,2)) for r in d['sweep']])"
