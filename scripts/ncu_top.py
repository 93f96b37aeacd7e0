"""Summary of an ncu report (raw page + source page): key throughput metrics and the instructions with the most
stall samples.  usage: python scripts/ncu_top.py report.ncu-rep [kernel index] [top n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else -1
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
data = rows[2:] if len(rows) > 2 and not rows[1][0].isdigit() else rows[1:]
vals = data[kidx]
want = ['Kernel Name', 'gpu__time_duration.sum', 'sm__cycles_elapsed.avg', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum',
        'SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg', 'SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg',
        'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct', 'dram__throughput.avg.pct_of_peak_sustained_elapsed']
for w in want:
    if w in hdr:
        print('  %-80s %s' % (w, vals[hdr.index(w)][:90]))
for h, v in zip(hdr, vals):
    if 'issue_stalled' in h and h.endswith('_per_issue_active.ratio') and float(v or 0) > 0.3:
        print('  %-80s %s' % (h.replace('smsp__average_warps_issue_stalled_', 'stall: ').replace('_per_issue_active.ratio', ''), v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
# the source page lists the kernels one after another: split on the "Kernel Name" lines
blocks, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == 'Kernel Name':
        cur = []
        blocks.append(cur)
    elif cur is not None:
        cur.append(r)
blk = blocks[kidx]
h2 = blk[0]
ix = {h: i for i, h in enumerate(h2)}
out, tot = [], 0
for n, r in enumerate(blk[1:]):
    if len(r) < len(h2):
        continue
    s = int(r[ix['# Samples']] or 0)
    tot += s
    out.append((s, n, r))
out.sort(reverse=True)
print('total samples', tot)
for s, n, r in out[:topn]:
    print('%5d %-50s %7d %5.1f%% ex %9s shw %9s/%-9s | L %s S %s mio %s math %s wait %s bar %s nsel %s' % (
        n, r[ix['Source']].strip()[:50], s, 100.0 * s / max(tot, 1), r[ix['Instructions Executed']], r[ix['L1 Wavefronts Shared']],
        r[ix['L1 Wavefronts Shared Ideal']], r[ix['stall_long_sb']], r[ix['stall_short_sb']], r[ix['stall_mio']], r[ix['stall_math']],
        r[ix['stall_wait']], r[ix['stall_barrier']], r[ix['stall_not_selected']]))
