# usage: bash tools/run_e2e_var.sh -- bench.py's end-to-end loop for different numbers of host stream groups
run() {
  python bench.py --steps 4 --warmup 2 --no-extra --no-cpu-baseline "$@" 2>&1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%-40s value %.0f e2e %.0f' % ('$*', d['value'], d['e2e']['value']))"
}
for g in 2 3 4 6; do run --e2e-groups $g; done
