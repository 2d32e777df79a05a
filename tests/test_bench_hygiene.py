"""bench.py on the CPU: what the two arms must agree on, and what the reference arm must not touch."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_problem_setup_maps_nothing_of_the_product():
    """`--impl reference` builds its problem through oracle/host_oracle.py only: neither the package nor its shared
    library may be loaded (the driver records which .so files each arm maps), and the meshes are the product's."""
    code = (
        "import sys, bench\n"
        "for name, n in (('config3', 408), ('config4', 4080), ('config5', 408)):\n"
        "    deck, mesh = bench.oracle_problem(name)\n"
        "    assert len(mesh[0]) == n and deck.energygroups == 4, (name, len(mesh[0]))\n"
        "assert 'nraps_b200' not in sys.modules and 'tests.util' not in sys.modules\n"
        "maps = open('/proc/self/maps').read()\n"
        "assert 'libnraps_b200' not in maps\n"
        "print('ok')\n"
    )
    run = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=120)
    assert run.returncode == 0 and run.stdout.strip() == "ok", run.stderr[-2000:]


def test_both_arms_print_the_same_config_object():
    sys.path.insert(0, ROOT)
    import bench

    for name in bench.WORKLOADS:
        for world in (1, 2, 8):
            a = bench.workload_config(name, world)
            b = bench.workload_config(name, world, "surface")
            assert a == b and set(a) == {"workload", "histories_per_generation", "source_mode", "tracking_mode"}
    assert bench.workload_config("config3", 8)["histories_per_generation"] == 80_000_000   # weak: 1e7 per GPU
    assert bench.workload_config("config4", 8)["histories_per_generation"] == 100_000_000  # strong: 1e8 in total
    assert bench.workload_config("config5", 8)["histories_per_generation"] == 1_000_000_000
