"""nraps_b200 -- B200-native Monte Carlo k-eigenvalue transport path of NRAPS.

Public surface mirrors the reference driver (src/main.rs:332-364):
``process_input`` -> ``mesh_gen`` -> ``monte_carlo`` -> ``plot_solution``.
All compute lives in ``nraps_b200/lib/libnraps_b200.so`` (CUDA, sm_100a), bound
through the C ABI in ``include/``; importing this package without that library
fails on first use.
"""
from .api import (  # noqa: F401
    DeltaX, Mesh, MonteCarloContext, SolutionResults, Variables, XSData, dev_div, dev_logf, dev_pcg32, format_f32, format_f64,
    make_options, mesh_gen, monte_carlo, nalgebra_method, plot_solution, process_input, trim,
)
from .dist import monte_carlo_distributed, shard_range  # noqa: F401

__all__ = [
    "DeltaX", "Mesh", "MonteCarloContext", "SolutionResults", "Variables", "XSData", "mesh_gen", "monte_carlo",
    "monte_carlo_distributed", "nalgebra_method", "plot_solution", "process_input", "shard_range",
]
