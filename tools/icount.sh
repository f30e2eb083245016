#!/bin/bash
# tools/icount.sh "<env>" [bench args]: warp instructions + duration of one frontier_kernel launch (ncu, 2 metrics only)
envs="$1"; shift
env $envs ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:frontier_ -s 3 -c 1 --csv python bench.py --ncu --steps 3 --warmup 3 "$@" 2>/dev/null | python -c "
import sys,csv
rows=[r for r in csv.reader(sys.stdin) if len(r)>5 and r[0].isdigit()]
d={r[-3]:r[-1] for r in rows}
print('$envs $*', {k:v for k,v in d.items()})"
