"""Golden vectors for the CSV layout: what the reference's own plot.py reads out of the files we write.

Run in the build container (needs /root/reference; the fixture travels, the reference does not):

    python tools/make_plot_reader_golden.py > tests/golden/plot_reader.json

For each case the host library writes vars.csv / k_eff.csv / interface.csv (nraps_plot_solution, the replacement
of src/plot_solution.rs:7-58) from a seeded SolutionResults, then the READING part of the reference's plot.py --
its source up to the first plotting statement, executed unmodified with matplotlib stubbed out -- parses them in
that directory.  The names it binds (length, meshed, generations, k, k_fund, flux0.., average0.., fission) go into
the fixture next to the file texts.  plot.py:27-37 is the live 2-group reader; the 4-group reader upstream keeps
commented out right below it (plot.py:39-56) is enabled for the G = 4 case by swapping the two blocks' comment
marks -- no other edit.  tests/test_host.py::test_reference_plot_reader_reads_our_csv checks the fixture against the
writer and the values.
"""
import json
import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nraps_b200 as nb  # noqa: E402

PLOT_PY = "/root/reference/plot.py"
f32 = np.float32


def results_for(G, N, gens, seed):
    """Seeded SolutionResults with the magnitudes of a real run (flux ~1e18, averages, a source with zero water cells)."""
    rng = np.random.default_rng(seed)
    return nb.SolutionResults(
        flux=(rng.random((G, N)) * 1e18).astype(f32), assembly_average=(rng.random((G, N)) * 1e18).astype(f32),
        fission_source=np.r_[rng.random(N - 2), 0, 0].astype(f32), k=(rng.random(gens) + 1).astype(f32),
        k_fund=np.r_[0, rng.random(gens - 1) + 1].astype(f32))


CASES = [dict(name="two_groups", G=2, N=11, gens=6, seed=11, length=float(f32(42.908089))),
         dict(name="four_groups", G=4, N=9, gens=5, seed=12, length=float(f32(41.598)))]


def reader_source(four_groups: bool) -> str:
    lines = open(PLOT_PY).read().split("\n")
    stop = next(i for i, l in enumerate(lines) if l.startswith("x = np.linspace"))
    lines = lines[:stop]
    if four_groups:
        # inside the interface.csv loop: comment the live if / elif chain, uncomment the 4-group chain below it
        start = next(i for i, l in enumerate(lines) if "interface.csv" in l) + 1
        live_end = next(i for i in range(start, len(lines)) if lines[i].strip() == "")
        dead_end = next(i for i in range(live_end + 1, len(lines)) if lines[i].strip() == "")
        for i in range(start, live_end):
            lines[i] = "    # " + lines[i][4:]
        for i in range(live_end + 1, dead_end):
            assert lines[i].startswith("    # "), lines[i]
            lines[i] = "    " + lines[i][6:]
    return "\n".join(lines) + "\n"


def main():
    plt_stub = types.ModuleType("matplotlib.pyplot")
    mpl_stub = types.ModuleType("matplotlib")
    mpl_stub.pyplot = plt_stub
    sys.modules.setdefault("matplotlib", mpl_stub)
    sys.modules.setdefault("matplotlib.pyplot", plt_stub)
    out = {"_comment": "made by tools/make_plot_reader_golden.py from /root/reference/plot.py (reader part, unmodified but for "
                       "the comment marks of its two interface.csv blocks in the 4-group case)", "cases": []}
    for c in CASES:
        r = results_for(c["G"], c["N"], c["gens"], c["seed"])
        with tempfile.TemporaryDirectory() as d:
            nb.plot_solution(r, c["G"], c["gens"], c["N"], c["length"], d)
            files = {n: open(os.path.join(d, n)).read() for n in ("vars.csv", "k_eff.csv", "interface.csv")}
            ns = {}
            cwd = os.getcwd()
            os.chdir(d)
            try:
                exec(compile(reader_source(c["G"] == 4), "plot.py(reader)", "exec"), ns)
            finally:
                os.chdir(cwd)
        read = {}
        for key, val in ns.items():
            if key in ("length", "meshed", "generations"):
                read[key] = val
            elif isinstance(val, np.ndarray):
                read[key] = [float(x) for x in val]
        out["cases"].append(dict(c, files=files, read=read))
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
