#!/bin/bash
# dev aid: run the MPC parity tests and a short bench with each prebuilt library variant in build_variants/ (GPU box scratch copy only)
cd "$(dirname "$0")/.."
LIB=hunter_bipedal_control_b200/libhunter_b200.so
cp $LIB /tmp/lib_default.so
for v in default "$@"; do
  if [ "$v" != default ]; then cp build_variants/$v $LIB; fi
  echo "=== variant $v"
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mpc or control or flow" 2>&1 | tail -3
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), {k: round(v,3) for k,v in d['kernel_ms_per_step'].items()})"
done
cp /tmp/lib_default.so $LIB
