#!/usr/bin/env python
"""Print the interesting parts of a bench.py JSON line (one file per argument)."""
import json
import sys

for f in sys.argv[1:]:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    det = d['detail']
    print(f"== {f}\nvalue {d['value']:.4f} s  e2e {d['e2e']['value'] if d.get('e2e') else None}  n_gpus {d['n_gpus']}  launches/step {d['gpu_launches'] / d['steps']:.0f}  "
          f"digest {d['result_digest'][:12]}  n_swaps {d['n_swaps']}")
    print("  parity", d.get('parity_spot'))
    print("  phases", det['phase_ms_last_step'], "\n  stats", det.get('lib_stats_rank0_last_step'))
    for r in det['roofline_all']:
        print(f"  {r['kernel'][:44]:44s} {r['achieved']:8.1f} {r['unit']:7s} frac {r['frac']:.3f} share {r['share_of_step']:.3f} launches {r['launches']} avg {r['avg_launch_ms']:.3f} ms")
        if 'by_mode' in r:
            for k, v in r['by_mode'].items():
                print(f"      {k:12s} launches {v['launches']:6d} ms {v['ms']:9.1f} gbs {v['gbs'] and round(v['gbs'], 1)}")
    for leg in ('sat_false', 'swap_heavy', 'cooperative', 'cfg5'):
        L = det.get(leg)
        if not L:
            continue
        if 'skipped' in L:
            print(f"  [{leg}] skipped: {L['skipped']}")
            continue
        print(f"  [{leg}] value {L['value']:.4f} digest {L['result_digest'][:12]} swaps/step {L['n_swaps_per_step']} launches/step {L['gpu_launches_per_step']:.0f} "
              f"phases {L['phase_ms_last_step']} stats {L.get('lib_stats_rank0_last_step')}")
        if L.get('parity_spot'):
            print("     parity", L['parity_spot'])
        for k, v in L.get('kernels', {}).items():
            print(f"     {k:20s} {v['achieved']:8.1f} {v['unit']:6s} frac {v['frac']:.3f} share {v['share_of_step']:.3f} launches {v['launches']}")
