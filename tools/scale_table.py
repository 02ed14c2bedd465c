#!/usr/bin/env python
"""Turn the bench lines of a strong-scaling run (profiles/<prefix>_n{1,2,4,8}.json, written by tools/gpu_scale2.sh)
into the markdown table of profiles/<prefix>.md.  `<prefix>_n4_rerun.json`, when present, is listed as an extra row.

    python tools/scale_table.py profiles/r02_scale_cfg5 > profiles/r02_scale_cfg5.md
"""
import json
import os
import sys


def line(path):
    text = open(path).read()
    return json.loads(text[text.index("{"):])


def main(prefix):
    rows = []
    for n in (1, 2, 4, 8):
        for tag in ("", "_rerun"):
            path = f"{prefix}_n{n}{tag}.json"
            if os.path.exists(path):
                rows.append((n, tag, line(path)))
    v1 = next(d["value"] for n, tag, d in rows if n == 1)
    e1 = next(d["e2e"]["value"] for n, tag, d in rows if n == 1)
    out = ["| N | step ms (device, max over ranks) | per-rank step ms | evals/s | efficiency v_N/(N v_1) | e2e ms | e2e evals/s | "
           "e2e speed-up | checksum | parity |", "|---|---|---|---|---|---|---|---|---|---|"]
    side = ["| N | cfg3 ragged through LPT: step ms / evals/s / sites max-over-mean | cfg2 one shard per rank (weak): step ms / evals/s |",
            "|---|---|---|"]
    for n, tag, d in rows:
        e = d["e2e"]
        per_rank = d.get("ms_per_step_per_rank")
        per_rank = " ".join(f"{x:.2f}" for x in per_rank) if per_rank else "-"
        e_ms = e.get("ms_per_step") or 1e3 * d["config"]["loci"] * 800 * 2048 / e["value"]
        par = d.get("parity", {})
        out.append(f"| {n}{' (rerun)' if tag else ''} | {d['ms_per_step']:.4f} | {per_rank} | {d['value']:.3e} | "
                   f"{d['value'] / (n * v1):.3f} | {e_ms:.1f} | {e['value']:.2e} | {e['value'] / e1:.2f} | "
                   f"{d.get('checksum', par.get('checksum', '-'))} | {'ok' if par.get('ok') else par or '-'} |")
        oc = d.get("other_configs", {})
        c3, c2 = oc.get("cfg3", {}), oc.get("cfg2_weak", oc.get("cfg2", {}))
        bal = (c3.get("partition") or {}).get("sites_per_rank_max_over_mean", "-")
        if c3 and c2:
            side.append(f"| {n}{' (rerun)' if tag else ''} | {c3['ms_per_step']:.4f} / {c3['value']:.3e} / {bal} | "
                        f"{c2['ms_per_step']:.4f} / {c2['value']:.3e} |")
    print("\n".join(out))
    print()
    print("Secondary workloads in the same runs (device-timed):\n")
    print("\n".join(side))


if __name__ == "__main__":
    main(sys.argv[1])
