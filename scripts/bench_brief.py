#!/usr/bin/env python
"""One line per configuration of a bench.py JSON line.  usage: bench_brief.py <file.json> ..."""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().split("\n")[-1])
    except Exception as e:
        print(f, "unreadable:", e); continue
    print(f"== {f}  n_gpus={d.get('n_gpus')} steps={d.get('steps')}")
    for r in [d] + d.get("extra_configs", []):
        if "roofline" not in r:
            print(f"  {r.get('name'):16s} value={r['value']:.4e} {r.get('unit')} ms={r.get('ms_per_step', r.get('ms_per_call', 0)):.4f} "
                  f"parity={r.get('parity_checked')} admit={r.get('admit_passes')}/{r.get('admit_skipped')} gb/s={r.get('gb_per_s')}")
            continue
        e = r.get("e2e") or {}
        print(f"  {r.get('name', 'headline'):9s} value={r['value']:.4e} ms/step={r['ms_per_step']:.4f} frac={r['roofline']['frac']:.4f} "
              f"e2e={e.get('value', 0):.4e} e2e_ms={e.get('ms_per_step', 0):.4f} parity={r.get('parity_checked')} "
              f"clk={r['clocks'].get('sm_mhz')} {r['clocks'].get('reasons')} subs/gpu={r['config']['subscribers_per_gpu']}")
