#!/bin/bash
dec() { timeout 300 python bench.py --workload decode --steps 3 --warmup 1 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), 'tok/s', round(d['ms_per_token_step'],4), 'ms/step frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value']))"; }
echo "== graph + PDL (default)"; dec
echo "== graph, no PDL"; B2_NO_PDL=1 dec
echo "== no graph, PDL"; B2_NO_GRAPH=1 dec
echo "== no graph, no PDL"; B2_NO_GRAPH=1 B2_NO_PDL=1 dec
