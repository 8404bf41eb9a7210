#!/bin/bash
# One gpurun call that produces everything profiles/ needs for a round (run as: gpurun --timeout 2400 -- 'bash tools/gpu_round_check.sh r2').
# Outputs land in gpurun_out/<tag>_*; copy the summaries into profiles/ afterwards (tools/summarize_launches.py, tools/ncu_summary.py,
# tools/ncu_lines.py).
set -u
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee $out/${tag}_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $out/${tag}_smoke.txt
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
python tools/lba_phases.py 72 2 > $out/${tag}_lba_phases_b72.txt 2>/dev/null
python tools/lba_phases.py 1 8 > $out/${tag}_lba_phases_b1.txt 2>/dev/null
# launch list of the device-resident loop (per-launch times are cold-cache and serialised: only the shares are comparable)
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 3 --rounds 3 --no-cpu-baseline --no-e2e --no-extra --dev-groups 1 > $out/${tag}_launches.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:lba_cluster -s 1 -c 1 -o $out/${tag}_ncu_lba \
    python bench.py --steps 1 --warmup 3 --rounds 3 --no-cpu-baseline --no-e2e --no-extra --dev-groups 1 > $out/${tag}_ncu_lba.log 2>&1
bash tools/ncu_frames.sh $tag > /dev/null 2>&1
tail -c 300 $out/${tag}_bench_n1.json
