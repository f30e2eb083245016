#!/bin/bash
# quick A/B: tools/q.sh "<env assignments>" [bench args]  -> one line: evals/s, ms/step, kernel_ms, frac
envs="$1"; shift
out=$(env $envs python bench.py --no-cpu --no-e2e --steps 120 --warmup 6 "$@" 2>&1 | tail -1)
python - "$envs $*" "$out" <<'PY'
import json,sys
try:
    d=json.loads(sys.argv[2]); r=d["roofline"]
    print("%-40s %.3e evals/s  %.4f ms/step  kernel %.4f ms  frac %.3f  grid %s smem %s st %s" % (sys.argv[1], d["value"], d["ms_per_step"], r["kernel_ms"], r["frac"], d["config"]["grid"], d["config"]["smem"], d["config"]["stages"]))
except Exception as e:
    print(sys.argv[1], "FAILED", sys.argv[2][-400:])
PY
