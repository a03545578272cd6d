"""Turn bench.py JSON lines (gpurun_out/*_bench_*.json) into the markdown rows used in profiles/README.md.

    python profiles/summarize_bench.py gpurun_out/r02a_bench_plain.json gpurun_out/r02a_bench_fused.json
"""
import json
import sys


def row(path):
    txt = open(path).read().strip().splitlines()
    d = json.loads(txt[-1])
    r, rp, rf = d.get('roofline') or {}, d.get('roofline_p2g_g2p') or {}, (d.get('roofline_fused') or d.get('roofline_g2p2g') or {})
    fb = d.get('fwd_bwd') or {}
    ring = (fb.get('whole_trajectory_ring') or {}).get('value')
    ob = (d.get('e2e_obs_bridge') or {}).get('value')
    f = lambda v, p=0: '—' if v is None else (f'{v:,.{p}f}')
    return (f"| `{path.split('/')[-1]}` | {d['n_gpus']} | {'fused' if d['config'].get('g2p2g_fused') else 'plain'} | {f(d['value'])} | {f(d['e2e']['value'])} | {f(ob)} | "
            f"{f(fb.get('value'))} | {f(ring)} | {f(r.get('launch_ms', None) and r['launch_ms'] * 1e3, 1)} | {f(r.get('frac'), 3)} | {f(rp.get('frac'), 3)} | "
            f"{f(rf.get('launch_ms', None) and rf['launch_ms'] * 1e3, 1)} | {f(rf.get('frac'), 3)} | {(d.get('clocks') or {}).get('sm_mhz')} |")


if __name__ == '__main__':
    print('| file | GPUs | path | substeps/s | e2e | e2e obs bridge | fwd+bwd pairs/s | … whole-trajectory ring | p2g µs | p2g frac | p2g+g2p frac | g2p2g µs | g2p2g frac | SM MHz |')
    print('|---|---|---|---|---|---|---|---|---|---|---|---|---|---|')
    for p in sys.argv[1:]:
        print(row(p))
