"""tests/golden/reference_shipped_shape.json from the reference's shipped interface.csv / k_eff.csv.

Those files are the only end-to-end artefacts the reference ships for the Monte Carlo path.  They come from an
older (f64) build whose results differ from HEAD's semantics (k = 5.67 = 3.11x, flux 2.9x in fuel to 4.2x in water), so only
scale-free quantities are kept: the fission-source shape, the thermal-flux shape and the group-mean flux ratios.
Usage (build container only): python tools/make_reference_shape.py
"""
import json
import os

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "reference_shipped_shape.json")

rows = [np.array(line.strip().split(","), dtype=float) for line in open(f"{REF}/interface.csv")]
flux, fission = np.array(rows[:4]), rows[8]
k = np.array(open(f"{REF}/k_eff.csv").readline().strip().split(","), dtype=float)
out = {
    "source": "reference interface.csv rows 1-4 (flux) and 9 (fission source), k_eff.csv row 1; TestCaseC-shaped run, older build",
    "n_cells": int(flux.shape[1]),
    "group_mean_flux_ratio": (flux.mean(axis=1) / flux.mean(axis=1)[0]).round(6).tolist(),
    "fission_source_shape": (fission / fission.sum()).round(9).tolist(),
    "thermal_flux_shape": (flux[3] / flux[3].sum()).round(9).tolist(),
    "k_relative_sd_per_generation": float(k.std(ddof=1) / k.mean()),
    # the four flux rows themselves (7 significant digits), for the piecewise-proportionality test: within one material
    # the shipped flux is a constant multiple of the flux of HEAD's semantics
    "flux_rows": [[float(f"{v:.7g}") for v in row] for row in flux],
    "generations": int(len(k)),
}
json.dump(out, open(OUT, "w"))
print("wrote", OUT)
