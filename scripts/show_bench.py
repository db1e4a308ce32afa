"""Compact view of a bench.py JSON line."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])


def kern(r):
    return ", ".join(f"{k['kernel']} {k['avg_ms']:.3f}ms x{k['launches_per_step']:g} {k['gbs']:.0f}GB/s" for k in r["kernels"])


r = d["roofline"]
print(f"headline {d['config']['workload']}: {d['ms_per_step']:.3f} ms/step {d['value'] / 1e9:.2f} G/s  clocks {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
print(f"  roofline {r['kernel']} {r['achieved']:.0f} GB/s frac {r['frac']:.3f} traffic {r.get('traffic')}; step model frac {r['step_model']['frac']:.3f}")
print("  " + kern(r))
print(f"  e2e {d['e2e'].get('ms_per_step', 0):.1f} ms/step {d['e2e']['value'] / 1e9:.3f} G/s; cpu {d.get('cpu_baseline', {}).get('value', 0) / 1e6:.2f} M/s; parity {d.get('parity', {}).get('rel_l2')}")
if "cfg2" in d:
    c = d["cfg2"]
    print(f"cfg2: {c['ms_per_step']:.4f} ms/step {c['value'] / 1e9:.2f} G/s frac {c['roofline']['frac']:.3f}; {kern(c['roofline'])}")
for w in d.get("workloads", []):
    r = w.get("roofline", {})
    print(f"{w['workload']}: {w.get('ms_per_step', 0):.3f} ms/step {w.get('value', 0) / 1e9:.2f} G/s path={w.get('path')} jit={w.get('specialised_kernels')} "
          f"contract frac {r.get('step_model', {}).get('frac', 0):.3f} moved {r.get('step_model', {}).get('bytes_per_cell_update_moved_by_this_path')}")
    if r:
        print("  " + kern(r))
