"""One-screen digest of bench.py JSON lines: python tools/bench_digest.py file.json [...]"""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable:", e)
        continue
    print(f"== {f}: N={d.get('n_gpus')} value {d['value']:.4g} ({d['ms_per_step']:.2f} ms/step)")
    if d.get("e2e"):
        e = d["e2e"]
        print(f"   e2e {e['value']:.4g} ({e.get('ms_per_step', 0):.2f} ms/step), d2h {e['d2h_bytes_per_step']/1e9:.3f} GB/rank, "
              f"ingest {e.get('host_ingest_gbs')}, bound {e.get('bound_ms', 0):.1f} ms, ms/bound {e.get('ms_over_bound', 0):.3f}")
    r = d.get("roofline")
    if r:
        print(f"   roofline frac {r['frac']:.3f} ({r['achieved']:.2f}/{r['peak']:.2f} TFLOP/s), kernel {r['kernel_ms_per_pass']:.2f} ms, "
              f"share {r['kernel_share_of_step']:.3f}, flop/step {r['flop_per_traj_step']}")
    if d.get("parity"):
        p = d["parity"]
        print("   parity", {k: v for k, v in p.items() if k not in ("checker", "subsample", "bars", "cross_rank_note")})
    if d.get("e2e_c5"):
        c = d["e2e_c5"]
        print(f"   c5 {c['value']:.4g} traj-steps/s, {c['ms_per_pass']:.1f} ms/pass, {c['n_trajectories_total']} trajectories, "
              f"impacted {c['impacted']}, alive {c['not_impacted_within_cap']}, nonfinite {c['nonfinite']}, checksum {c['checksum_impact_time']:.6f}")
    if d.get("cpu_baseline"):
        print("   cpu", json.dumps(d["cpu_baseline"])[:400])
    print("   clocks", d.get("clocks"))
