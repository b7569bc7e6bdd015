timeout 300 python tools/_diag.py 2>&1 | tail -4
