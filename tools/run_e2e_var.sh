# usage: bash tools/run_e2e_var.sh -- sweeps bench.py's end-to-end loop over host-side scheduling options
run() {
  env "$1" python bench.py --steps 4 --warmup 2 --no-extra --no-cpu-baseline "${@:2}" 2>&1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%-60s value %.0f e2e %s' % ('$*', d['value'], d.get('e2e')))"
}
run X=1 --e2e-groups 4
run X=1 --e2e-groups 8
run BENCH_SWITCH_INTERVAL=5e-3 --e2e-groups 4
run BENCH_E2E_LBA=exclusive --e2e-groups 4
