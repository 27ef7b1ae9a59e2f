#!/bin/bash
# sweep the tuning knobs of the TMA local-operator kernel (needs a build with HB_EXTRA_NVCC_FLAGS=-DHB_TUNE_SHAPES)
export PYTHONPATH=.
python tools/repro_local.py || exit 1
for shape in 0 1 2 3 4; do for promo in 0 2 3; do for ctas in 0 2 3 4; do
  echo -n "shape=$shape promo=$promo ctas=$ctas: "
  HB_LOCAL_TMA_SHAPE=$shape HB_TMA_L2PROMO=$promo HB_TMA_CTAS=$ctas python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['roofline']['frac'],3))"
done; done; done
