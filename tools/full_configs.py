"""BASELINE configs 3 and 5 at their full generation counts on one GPU: k_fund, sigma, entropy (documentation run)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nraps_b200 as nb
from tests.util import load_case

args = load_case("c")
out = {}
for name, kw in [("config3_surface", dict(generations=200, histories=10_000_000)),
                 ("config3_woodcock", dict(generations=200, histories=10_000_000, tracking_mode="woodcock")),
                 ("config5_1gpu_surface_bank", dict(generations=20, histories=125_000_000, source_mode="fission_bank")),
                 ("config5_1gpu_woodcock_bank", dict(generations=20, histories=125_000_000, source_mode="fission_bank", tracking_mode="woodcock"))]:
    t0 = time.time()
    r = nb.monte_carlo(*args, 1.0, skip=5 if "bank" in name else 1, **kw)
    sk = 5 if "bank" in name else 1
    k = r.k[sk:].astype(np.float64)
    out[name] = dict(k_fund=float(r.k_fund[-1]), k_mean=float(k.mean()), sigma_mean=float(k.std(ddof=1) / np.sqrt(len(k))),
                     sigma_generation=float(k.std(ddof=1)), device_s=r.seconds_device, wall_s=time.time() - t0,
                     histories_per_s=kw["generations"] * kw["histories"] / r.seconds_device,
                     collisions_per_history=r.counters["collisions"] / r.counters["histories"],
                     entropy_first_last=[float(r.entropy[0]), float(r.entropy[-1])], bank_last=int(r.bank_sizes[-1]),
                     flux_rel_std_error_median=float(np.median(r.flux_std_error(kw["generations"], sk) / r.flux)))
    print(name, json.dumps(out[name]), flush=True)
json.dump(out, open("gpurun_out/full_configs.json", "w"), indent=1)
