#!/bin/bash
# brief bench line: infer value, e2e, train value, train ms, roofline TFLOP/s
python bench.py --steps 20 --warmup 5 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); t=d.get('train') or {}; print('infer %.1f e2e %.1f train %.1f (%.2f ms) fc1 %.0f TF/s clocks %s' % (d['value'], d['e2e']['value'], t.get('value',0), t.get('ms_per_step',0), d['roofline']['achieved'], d['clocks']))"
