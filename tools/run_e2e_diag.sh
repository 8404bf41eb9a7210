# usage: bash tools/run_e2e_diag.sh -- which half of the end-to-end loop costs what (diagnostic runs; the lines are marked invalid)
run() { env "$1" python bench.py --steps 4 --warmup 2 --no-extra --no-cpu-baseline "${@:2}" 2>&1 | tail -1 | cut -c1-300; }
run BENCH_E2E_DIAG=nolba --e2e-groups 2
run BENCH_E2E_DIAG=nolba --e2e-groups 4
run BENCH_E2E_DIAG=nolba --e2e-groups 8
run BENCH_E2E_DIAG=noframes --e2e-groups 4
