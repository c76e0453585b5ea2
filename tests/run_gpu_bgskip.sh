SMALL="python bench.py --steps 4 --warmup 3 --pool-cards 256 --pool-bgs 128 --no-e2e --no-cpu-baseline"
for v in "" $VARIANTS; do
  if [ -n "$v" ]; then export MTGV_LIB=$PWD/mtgvision_b200/csrc/variants/libmtgv_$v.so; else unset MTGV_LIB; fi
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_background -c 6 --csv $SMALL 2>/dev/null | grep k_background | tail -3 | awk -F'","' -v v="$v" '{s+=$NF} END {print "variant=[" v "] k_background us:", s/3/1000}' | tr -d '"'
done
