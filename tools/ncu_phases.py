"""Per-phase share of executed instructions and stall samples for the gp64 kernels.
   python tools/ncu_phases.py gpurun_out/prof.ncu-rep"""
import csv, io, subprocess, sys
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur = None; h = None; f = None; data = {}; stall = {}
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': f = r[1].split('/')[-1]; continue
    if len(r) >= 2 and r[0] == 'Function Name': cur = r[1]; data.setdefault(cur, []); continue
    if r and r[0] == 'Line No': h = r; continue
    if cur and h and len(r) == len(h) and r[0].isdigit():
        st = {n[6:]: int(r[i]) for i, n in enumerate(h) if n.startswith('stall_') and 'Not Issued' not in n}
        data[cur].append((f, int(r[0]), int(r[h.index('Instructions Executed')]), int(r[h.index('# Samples')]), st))
srcl = open('cosmogp_b200/csrc/cgp_small64.cu').read().split('\n')
def find(pat):
    for i, l in enumerate(srcl):
        if pat in l: return i + 1
marks = [('stage', find('---------------- stage')), ('K', find('phase K: covariance')), ('C-update', find('phase C: left-looking')),
         ('C-sub', find('// C = K (parked) - update')), ('diag', find('diag_factor(s0[0]')), ('park/trsm', find('park C[I][J] to re-read')),
         ('LLsolve', find('z = L^-1 r by block forward')), ('Linv', find('L^-1 in place, row by row')),
         ('zalpha', find('z = L^-1 r (and L^-1 1)')), ('loo', find('const double rho = cov.amp_cross')),
         ('predict-setup', find('two blocks of 8 grid points per pass')), ('predict-exp', find('cross-covariance fragments (no amplitude)')),
         ('predict-dmma', find('double acc0[2][NB], acc1[2][NB];')), ('predict-out', find('double vv = 0.0, vv2 = 0.0;'))]
marks = [m for m in marks if m[1]]
d0, d1 = find('__device__ __forceinline__ void diag_factor'), find('constexpr int n_vec64')
def phase(f, l):
    if f == 'cgp_math.cuh': return 'rsqrt' if l >= 45 else 'EXP'
    if f != 'cgp_small64.cu': return 'shfl/sync'
    if d0 <= l < d1 - 12: return 'diag'
    if l < marks[0][1]:
        return 'dmma/ld helpers'
    p = 'pre'
    for name, ln in marks:
        if l >= ln: p = name
    return p
for k, v in data.items():
    tot = sum(x[2] for x in v); ts = sum(x[3] for x in v)
    agg = {}
    for f, l, i, s, st in v:
        a = agg.setdefault(phase(f, l), [0, 0, {}]); a[0] += i; a[1] += s
        for n, c in st.items(): a[2][n] = a[2].get(n, 0) + c
    print(k[40:78], 'inst/obj %.0f' % (tot / 1e5))
    for p, (i, s, st) in sorted(agg.items(), key=lambda x: -x[1][1]):
        top = ', '.join('%s %.0f%%' % (n, 100 * c / max(1, sum(st.values()))) for n, c in sorted(st.items(), key=lambda x: -x[1])[:4])
        print('   %-16s inst %5.1f%% (%6.0f/obj)  time %5.1f%%   [%s]' % (p, 100 * i / tot, i / 1e5, 100 * s / ts, top))
