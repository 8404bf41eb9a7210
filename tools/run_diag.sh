for lib in "" build/liborb_lba256.so; do
for diag in nolba noframes; do
  for g in 1 2 4; do
    echo -n "lib=$lib diag=$diag groups=$g: "; env ${lib:+ORB_B200_LIB=$lib} BENCH_DIAG=$diag python bench.py --steps 6 --warmup 2 --dev-groups $g 2>&1 | tail -1
  done
done
done
