"""Weak-scaling table from committed bench lines: python tools/scaling_summary.py profiles/bench_r2p_n1.json profiles/bench_r2j_n4_peer.json ..."""
import json, sys
rows = []
for path in sys.argv[1:]:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    rows.append((d["n_gpus"], path, d))
rows.sort(key=lambda r: (r[0], r[1]))
base = next((d for n, _, d in rows if n == 1 and d["config"].get("per_gpu_block") == rows[0][2]["config"].get("per_gpu_block")), None)
print(f"{'file':44s} {'N':>2s} {'transport':>9s} {'ms/iter':>8s} {'it/s':>8s} {'value':>9s} {'T1/TN':>6s} {'e2e/val':>7s} {'parity':>9s} {'n_global':>10s}")
for n, path, d in rows:
    t1 = base["ms_per_step"] if base and base["config"].get("per_gpu_block") == d["config"].get("per_gpu_block") else None
    eff = f"{t1 / d['ms_per_step']:.3f}" if t1 else "-"
    par = d["parity"].get("max_rel_err")
    print(f"{path.split('/')[-1]:44s} {n:2d} {str(d['details'].get('transport')):>9s} {d['ms_per_step']:8.4f} {d['iterations_per_s']:8.1f} {d['value']:9.1f} {eff:>6s} "
          f"{d['e2e']['fraction_of_value']:7.3f} {par if par is None else format(par, '.1e'):>9} {d['details']['n_global']:10d}")
