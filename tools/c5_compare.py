"""Compares the bench lines of the 1000-frame workload on 1 and on N GPUs (tools/c5_scaling.sh): per-iteration losses of the
frame-window sharded run against the single-GPU run, the aggregate rate ratio and the per-GPU roofline fractions.

    python tools/c5_compare.py gpurun_out/c5_n1.json gpurun_out/c5_n8.json
"""
import json
import sys


def last_json(path):
    with open(path) as f:
        rows = [ln for ln in f.read().splitlines() if ln.strip().startswith("{")]
    return json.loads(rows[-1])


a, b = last_json(sys.argv[1]), last_json(sys.argv[2])
ha, hb = a["loss_hist_head"], b["loss_hist_head"]
n = min(len(ha), len(hb))
rel = [abs(x - y) / abs(x) for x, y in zip(ha[:n], hb[:n])]
print(f"{a['n_gpus']} GPU: {a['value']:.2f} it/s ({a['ms_per_step']:.3f} ms), scan roofline {a['roofline']['frac']:.3f}, iteration {a['roofline_iteration']['frac']:.3f}")
print(f"{b['n_gpus']} GPU: {b['value']:.2f} it/s ({b['ms_per_step']:.3f} ms), scan roofline {b['roofline']['frac']:.3f} (rank 0), iteration {b['roofline_iteration']['frac']:.3f} per GPU")
print(f"rate ratio {b['value'] / a['value']:.2f}x on {b['n_gpus']} GPUs")
print(f"per-iteration loss, first {n} iterations: max relative difference {max(rel):.2e}")
for k in range(n):
    print(f"  it {k:2d}  {ha[k]:.8f}  {hb[k]:.8f}  {rel[k]:.1e}")
sys.exit(0 if max(rel[:10]) <= 1e-6 else 1)
