#!/usr/bin/env python
"""tools/show_bench.py FILE [KEY ...] -- pretty-print (part of) a bench.py JSON line."""
import json
import sys


def show(x, ind=0):
    for k, v in x.items():
        if isinstance(v, dict):
            print(" " * ind + f"{k}:")
            show(v, ind + 2)
        else:
            print(" " * ind + f"{k}: {str(v)[:160]}")


d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
for k in sys.argv[2:]:
    d = d[k]
show(d) if isinstance(d, dict) else print(d)
