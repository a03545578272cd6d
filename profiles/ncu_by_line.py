"""Dynamic instruction counts and stall samples PER SOURCE LINE of one kernel: joins the SASS-level `ncu --page source --csv` view of an
.ncu-rep (per-instruction 'Instructions Executed' / '# Samples') with the line table of the library the profile was taken from
(nvdisasm --print-line-info; the library is built with -lineinfo).  CPU-side tool, run in the build container:
    python profiles/ncu_by_line.py gpurun_out/x.ncu-rep fluidlab_b200/libfluidmpm.so k_fwd [n_warps]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, lib, pat = sys.argv[1:4]
nwarps = float(sys.argv[4]) if len(sys.argv) > 4 else None
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kname = rows[0][1]
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
base = int(data[0][ix['Address']], 16)
dyn = {int(r[ix['Address']], 16) - base: (int(r[ix['Instructions Executed']]), int(r[ix['# Samples']]), r[ix['Source']]) for r in data}
# line table of the same kernel
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, capture_output=True)
mang = None
lines = {}
for f in os.listdir(tmp):
    if not f.endswith('.cubin') or 'sm_100' not in f:
        continue
    txt = subprocess.run(['nvdisasm', '-c', '--print-line-info', os.path.join(tmp, f)], capture_output=True, text=True).stdout
    for sec in re.split(r'\n\s*//-+ \.text\.', txt)[1:]:
        name = sec.split()[0]
        dem = subprocess.run(['c++filt', name], capture_output=True, text=True).stdout.strip()
        if dem.replace(' ', '') .startswith(kname.replace(' ', '')[:len(dem.replace(' ', ''))]) or kname.replace('(bool)', '').replace('(int)', '').replace(' ', '').startswith(dem.replace('true', '1').replace('false', '0').replace(' ', '')[:20]) and pat in name:
            pass
        if pat not in name:
            continue
        # choose the section whose instruction count matches the profile
        cur = None; tab = {}
        for line in sec.split('\n'):
            m = re.search(r'//## File "([^"]+)", line (\d+)', line)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
            m = re.search(r'/\*([0-9a-f]{4,6})\*/\s+\S', line)
            if m:
                tab[int(m.group(1), 16)] = cur
        if len(tab) == len(dyn):
            lines[name] = tab
if not lines:
    sys.exit(f'no section matching {pat} with {len(dyn)} instructions')
name, tab = sorted(lines.items())[0]
if len(lines) > 1:
    print('# several candidate sections with the same length:', list(lines), file=sys.stderr)
agg = collections.defaultdict(lambda: [0, 0])
for off, (n, s, _) in dyn.items():
    k = tab.get(off)
    agg[k][0] += n; agg[k][1] += s
tot_n = sum(v[0] for v in agg.values()); tot_s = sum(v[1] for v in agg.values())
print(f'# {kname}\n# section {name}: {len(dyn)} SASS instructions, {tot_n} executed, {tot_s} stall samples')
src_cache = {}
def src(k):
    if k is None: return ''
    fn, ln = k
    for d in ('fluidlab_b200/csrc', 'include'):
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), d, fn)
        if os.path.exists(p):
            if p not in src_cache: src_cache[p] = open(p).read().split('\n')
            return src_cache[p][ln - 1].strip()[:110]
    return ''
for k, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(os.environ.get('TOP', 70))]:
    per = f'{n / nwarps:8.1f}/warp' if nwarps else ''
    print(f'{100 * n / tot_n:5.1f}% inst {per} {100 * s / max(tot_s, 1):5.1f}% stall  {k}  {src(k)}')
