SMALL="python bench.py --steps 30 --warmup 5 --pool-cards 512 --pool-bgs 256 --no-e2e --no-cpu-baseline"
for v in "" $VARIANTS; do
  if [ -n "$v" ]; then export MTGV_LIB=$PWD/mtgvision_b200/csrc/variants/libmtgv_$v.so; else unset MTGV_LIB; fi
  echo "variant=[$v]"; $SMALL | python -c "import json,sys; d=json.load(sys.stdin); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])"
done
