#!/bin/bash
# ab_env.sh <tag> <reps> "<ENV=a ...>" "<ENV=b ...>" ...: the default library under different environments, alternated
# `reps` times on the same box (device-resident step, staged step, stage times, end to end)
cd "$(dirname "$0")/../.."
TAG=$1; REPS=$2; shift; shift
for R in $(seq 1 $REPS); do
  I=0
  for E in "$@"; do
    I=$((I+1))
    env $E python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-api --other-configs none $BENCH_ARGS \
        > gpurun_out/ab_${TAG}_${I}_$R.json 2> gpurun_out/ab_${TAG}_${I}_$R.err
    python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/ab_${TAG}_${I}_$R.json').read().strip().split('\n')[-1])
    s = d['stage_ms_per_step']
    print('%-28s step %.4f  staged %.4f  lattice %.4f  beam %.4f  pack %.4f  e2e %.4f  sha1 %s' % ('$E', d['ms_per_step'], d.get('ms_per_step_staged', 0), s['ms_lattice'], s['ms_beam'], s['ms_pack'], d['e2e']['ms_per_step'], d['results_sha1'][:12]))
except Exception as e:
    print('$E failed', e)
PY
  done
done
