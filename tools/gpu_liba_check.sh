#!/bin/bash
# One gpurun call for the LocalInertialBA kernel: its parity tests first (default library and any thread-count variants under build/), timing, then the
# whole GPU suite, smoke, memcheck + racecheck, one ncu capture, the bench line.
set -u
tag=${1:-r2b}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_local_inertial_ba_gpu.py -q 2>&1 | tail -25 | tee $out/${tag}_pytest_liba.txt
timeout 300 python tools/liba_time.py 2>&1 | tail -5 | tee $out/${tag}_liba_time.txt
for v in nt256 nt512 nt128; do
  if [ -f build/liborb_$v.so ]; then
    echo "== $v" | tee -a $out/${tag}_liba_time.txt $out/${tag}_pytest_liba.txt
    ORB_B200_LIB=build/liborb_$v.so timeout 300 python -m pytest tests/test_local_inertial_ba_gpu.py -q 2>&1 | tail -5 | tee -a $out/${tag}_pytest_liba.txt
    ORB_B200_LIB=build/liborb_$v.so timeout 300 python tools/liba_time.py 2>&1 | tail -5 | tee -a $out/${tag}_liba_time.txt
  fi
done
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee $out/${tag}_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $out/${tag}_smoke.txt
timeout 300 compute-sanitizer --tool memcheck python tools/liba_ncu_target.py 2 small 2>&1 | tail -4 | tee $out/${tag}_sanitizer_memcheck_liba.txt
timeout 300 compute-sanitizer --tool racecheck python tools/liba_ncu_target.py 2 small 2>&1 | tail -4 | tee $out/${tag}_sanitizer_racecheck_liba.txt
timeout 400 ncu --set full --import-source on --clock-control none -k regex:local_inertial_ba -c 1 -o $out/${tag}_ncu_liba python tools/liba_ncu_target.py 37 > $out/${tag}_ncu_liba.log 2>&1
timeout 900 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
tail -c 1500 $out/${tag}_bench_n1.json
