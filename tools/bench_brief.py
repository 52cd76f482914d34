#!/usr/bin/env python3
"""print the headline fields of bench.py JSON lines found in the given log files"""
import json, sys
for f in sys.argv[1:]:
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l); r = d.get('roofline') or {}
            print(f, '%.3e' % d['value'], 'ms/mc', round(d['config'].get('ms_per_mc_step', 0), 4), 'e2e %.3e' % d['e2e']['value'],
                  {k: v for k, v in list((r.get('kernels_ms_per_mc_step') or {}).items())[:6]})
