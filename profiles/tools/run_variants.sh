#!/bin/bash
# run_variants.sh <tag> [config] [extra bench args]: times every lattice_based_tagger_b200/variants/*.so on the GPU box
# (device-resident step, stage times, digest of the results) — experiments; the default library is restored at the end.
cd "$(dirname "$0")/../.."
TAG=${1:-var}; CONFIG=${2:-c2}; shift; shift
LIB=lattice_based_tagger_b200/liblt_b200.so
cp $LIB /tmp/liblt_default.so
for V in lattice_based_tagger_b200/variants/*.so; do
    N=$(basename $V .so)
    cp $V $LIB
    python bench.py --config $CONFIG --steps 20 --warmup 3 --no-cpu-baseline --no-api --other-configs none "$@" \
        > gpurun_out/var_${TAG}_${CONFIG}_$N.json 2> gpurun_out/var_${TAG}_${CONFIG}_$N.err
    python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/var_${TAG}_${CONFIG}_$N.json').read().strip().split('\n')[-1])
    s = d['stage_ms_per_step']
    print('%-24s step %.4f  lattice %.4f  beam %.4f  e2e %.4f  sha1 %s' % ('$N', d['ms_per_step'], s['ms_lattice'], s['ms_beam'], d['e2e']['ms_per_step'], d['results_sha1'][:12]))
except Exception as e:
    print('$N failed', e)
PY
done
cp /tmp/liblt_default.so $LIB
