# usage: bash tools/run_e2e_var.sh -- sweeps bench.py's end-to-end loop over host-side scheduling options
run() {
  env $1 python bench.py --steps 4 --warmup 2 --no-extra --no-cpu-baseline "${@:2}" 2>&1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%-60s value %.0f e2e %.0f' % ('$*', d['value'], d['e2e']['value']))"
}
run "X=1"
run "BENCH_E2E_SPLIT=2"
run "BENCH_E2E_SPLIT=2 BENCH_LBA_PRIO=1"
run "BENCH_E2E_SPLIT=3 BENCH_LBA_PRIO=1"
run "BENCH_E2E_SPLIT=2 BENCH_LBA_PRIO=1" --e2e-groups 3
