#!/bin/bash
# A/B of bench.py variants on ONE box (boxes differ by ~1-2%): each variant in $VARIANTS (separated by ';') is run
# $REPS times, interleaved.  Logs to gpurun_out/ab_*.json.
mkdir -p gpurun_out
IFS=';' read -ra V <<< "${VARIANTS:-;--no-overlap}"
for rep in $(seq 1 ${REPS:-2}); do
  for i in "${!V[@]}"; do
    timeout 300 python bench.py --no-cpu-baseline --no-aux --steps ${STEPS:-100} --warmup 10 ${V[$i]} > gpurun_out/ab_${i}_${rep}.json 2> gpurun_out/ab_${i}_${rep}.err
    python - "$i" "$rep" "${V[$i]}" <<'PY'
import json, sys
i, rep, v = sys.argv[1:4]
try:
    d = json.loads(open('gpurun_out/ab_%s_%s.json' % (i, rep)).read().strip().splitlines()[-1])
    print("[%s] rep %s: %.4f ms/step %.0f pl/s | e2e %.4f ms | roof %.0f GB/s %.3f ms | %s" % (v or "default", rep, d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['roofline']['achieved'], d['roofline']['launch_ms'], {k: round(x, 3) for k, x in d['phase_ms'].items() if k in ('decode_loss_dz', 'dh', 'dw_dec', 'adam_dec', 'adam_enc', 'encode_fwd', 'ybits')}))
except Exception as e:
    print("[%s] rep %s failed: %s" % (v, rep, e)); print(open('gpurun_out/ab_%s_%s.err' % (i, rep)).read()[-1500:])
PY
  done
done
