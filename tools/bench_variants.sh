#!/bin/bash
# Kernel tuning helper (GPU box): time bench.py stages for every alternative build under build/variants/.
for so in intmax_zkp_core_b200/libb200zkp.so build/variants/*.so; do
  [ -f "$so" ] || continue
  B200ZKP_LIB=$PWD/$so B200ZKP_SKIP_CPU=1 timeout 300 python bench.py --steps ${STEPS:-2} --warmup 2 2>&1 | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$so', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['stages_ms_per_step'].items()}, "")"
done
