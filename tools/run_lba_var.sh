# usage: bash tools/run_lba_var.sh  -- sweeps bench.py's device-resident loop over scheduling options
run() {
  python bench.py --steps 6 --warmup 2 --no-e2e --no-extra --no-cpu-baseline "$@" 2>&1 | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%-44s value %.0f ms/round %.3f lba/round %.3f cluster %s' % ('$*', d['value'], d['ms_per_step'] / d['config']['rounds_per_step'], d['stage_ms_per_round']['lba_cluster_kernel'], d['config']['lba_cluster_size']))"
}
run --lba-rounds 3
run --lba-rounds 3 --dev-groups 2
run --lba-rounds 6
run --lba-rounds 3 --lba-concurrent
