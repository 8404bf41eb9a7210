"""Per-source-line warp-stall samples of one kernel in an ncu report (needs -lineinfo and --import-source on).
usage: python tools/ncu_lines.py report.ncu-rep [top_n]"""
import sys, collections
sys.path.insert(0, '/opt/nvidia/nsight-compute/2025.2.1/extras/python')
import ncu_report
ctx = ncu_report.load_report(sys.argv[1])
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
act = ctx.range_by_idx(0).action_by_idx(0)
samp = act.metric_by_name('smsp__pcsamp_sample_buffer_full') 
m = act.metric_by_name('smsp__pcsamp_warps_issue_stalled_total') if 'smsp__pcsamp_warps_issue_stalled_total' in act.metric_names() else None
names = [n for n in act.metric_names() if n.startswith('smsp__pcsamp_warps_issue_stalled_') and not n.endswith('_not_issued')]
inst = act.metric_by_name('inst_executed')
pcs = act.metric_by_name('smsp__pcsamp_sample_buffer_full')
lines = collections.Counter(); reasons = collections.defaultdict(collections.Counter); iex = collections.Counter()
for n in names:
    mm = act.metric_by_name(n)
    if mm.num_instances() == 0: continue
    cids = mm.correlation_ids()
    for i in range(mm.num_instances()):
        pc = cids.as_uint64(i); v = mm.as_uint64(i)
        if not v: continue
        si = act.source_info(pc)
        key = (si.file_name().split('/')[-1], si.line()) if si else ('?', 0)
        lines[key] += v; reasons[key][n.replace('smsp__pcsamp_warps_issue_stalled_', '')] += v
cids = inst.correlation_ids()
for i in range(inst.num_instances()):
    si = act.source_info(cids.as_uint64(i))
    key = (si.file_name().split('/')[-1], si.line()) if si else ('?', 0)
    iex[key] += inst.as_uint64(i)
tot = sum(lines.values())
print('total samples', tot, ' total warp-instr', sum(iex.values()))
src = {}
for (f, l), v in lines.most_common(topn):
    if f not in src:
        try: src[f] = open('/root/repo/orb_slam3_modified_b200/csrc/' + f).read().split('\n')
        except OSError: src[f] = []
    text = src[f][l - 1].strip()[:90] if 0 < l <= len(src[f]) else ''
    top = ', '.join('%s %d%%' % (k, 100 * c // v) for k, c in reasons[(f, l)].most_common(3))
    print('%5.1f%%  inst %8d  %s:%d  [%s]  %s' % (100.0 * v / tot, iex[(f, l)], f, l, top, text))
