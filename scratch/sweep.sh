#!/bin/bash
# diagnostics: step time of the headline bench under scheduling knobs (each line: knobs, ms/step, e2e ms/step)
run() { env "$@" python bench.py --steps 200 --warmup 10 --skip-cpu --windows 9 > /tmp/b.json 2>/tmp/b.err || tail -3 /tmp/b.err; python -c "
import json; d=json.load(open('/tmp/b.json')); print('$*', round(d['ms_per_step']*1e3,2), round(d['e2e']['ms_per_step']*1e3,2))"; }
run GS_DENSE_X1=0
run GS_DENSE_X1=1
run GS_AGG_GRID=rows
run GS_AGG_GRID=rows GS_TRAIN_PRIO=1
run GS_TRAIN_PRIO=1
run GS_BG_AGG_CTAS=2
run GS_BG_AGG_CTAS=2 GS_TRAIN_PRIO=1
