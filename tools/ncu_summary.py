"""Summarise an .ncu-rep: per kernel headline metrics, stall mix and the hottest source lines.
   python tools/ncu_summary.py gpurun_out/prof.ncu-rep [n_lines]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__grid_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.sum', 'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.max']
for r in rows[2:]:
    print('-----')
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print('  %s = %s %s' % (w, r[i], units[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur = None; data = {}; h = None; f = None; agg = {}
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': f = r[1].split('/')[-1]; continue
    if len(r) >= 2 and r[0] == 'Function Name': cur = r[1]; data.setdefault(cur, []); agg.setdefault(cur, {}); continue
    if r and r[0] == 'Line No': h = r; continue
    if cur and h and len(r) == len(h) and r[0].isdigit():
        data[cur].append([f] + r)
        for i, name in enumerate(h):
            if name.startswith('stall_') and 'Not Issued' not in name:
                agg[cur][name] = agg[cur].get(name, 0) + int(r[i])
ii = h.index('Instructions Executed') + 1; si = h.index('# Samples') + 1
for k, v in data.items():
    print('=====', k[:100])
    tot = sum(int(r[ii]) for r in v); tots = sum(int(r[si]) for r in v)
    print('total inst', tot, 'samples', tots)
    st = agg[k]; ts = sum(st.values()) or 1
    print('stalls:', ', '.join('%s %.1f%%' % (a[6:], 100 * b / ts) for a, b in sorted(st.items(), key=lambda x: -x[1])[:8]))
    for r in sorted(v, key=lambda r: -int(r[si]))[:nl]:
        print('%-20s %5s inst %5.1f%% samp %5.1f%%  %s' % (r[0][:20], r[1], 100 * int(r[ii]) / tot, 100 * int(r[si]) / tots, r[2][:105]))
