#!/bin/bash
set -u
for S in 4 8 12 16 24 32; do python bench.py --no-cpu --no-sharded --steps 10 --warmup 3 --streams $S 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('streams $S', round(d['value'],1), round(d['e2e']['value'],1))"; done
