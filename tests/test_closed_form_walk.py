"""The closed-form strides and the sure loop of the surface kernel (mc_transport.cu: skip_cells, the stride rounds, the
five-instruction loop over surely-crossed cells) on the CPU.

tools/closed_form_walk.c restates that arithmetic in plain C -- the mantissa recurrence of fl(ds - w) inside one
binade, the rounding-tie rule, the sure-crossing limit, the single real subtraction between strides, the power-of-two cut
of the strides, the test-free loop over the cells whose |ds| is above the limit -- and walks
random neutrons through the segments of six meshes (MPFR = 8 ... 640, edges accumulated in f32 like mesh_gen) both
ways: cell by cell as the reference does (src/mc_code.rs:151-181) and the kernel's way; a quarter of the walks start with
|ds| within 10 % of the sure-crossing margin or within an ulp or two of a whole number of cell widths.  Cell, ds bits, position bits,
collision flag and collision position must agree on every walk.  The GPU parity tests on the fine meshes check the
kernel itself; this one pins the arithmetic where a GPU is not needed."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_strides_equal_the_cell_by_cell_walk(tmp_path):
    exe = tmp_path / "closed_form_walk"
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-o", str(exe), os.path.join(ROOT, "tools", "closed_form_walk.c"), "-lm"],
                   check=True, capture_output=True)
    for seed in ("1", "77"):
        run = subprocess.run([str(exe), "150000", seed], capture_output=True, text=True, timeout=300)
        assert run.returncode == 0, run.stdout[-2000:]
        last = run.stdout.strip().splitlines()[-1]
        assert "bad so far 0" in last and "skipped" in last, last
