#!/bin/bash
# A/B of decoder variants (tests/build_variant.sh): decode timing of the e2e batch mix (MODES: rgb and/or pools) + JPEG parity tests
for v in $VARIANTS; do
  export MTGV_LIB=$PWD/mtgvision_b200/csrc/variants/libmtgv_$v.so
  echo "variant=[$v]"
  for m in ${MODES:-rgb pools}; do timeout 300 python tests/microbench/decode_mix.py 20 $m; done
  timeout 600 python -m pytest tests/test_gpu_jpeg.py -m gpu -x -q --timeout 300 2>&1 | tail -1
done
