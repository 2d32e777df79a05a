import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time, torch, nraps_b200 as nb
from tests.util import load_case
args = load_case("c")
torch.cuda.init(); torch.zeros(1, device="cuda"); torch.cuda.synchronize()
for rep in range(6):
    t0=time.perf_counter()
    ctx = nb.MonteCarloContext(*args, 1.0, generations=20, histories=10_000_000, skip=1)
    t1=time.perf_counter()
    ctx.transport(0); torch.cuda.synchronize(); t2=time.perf_counter()
    ctx.finalize_generation(0); torch.cuda.synchronize(); t3=time.perf_counter()
    for g in range(1,20):
        ctx.transport(g); ctx.finalize_generation(g)
    torch.cuda.synchronize(); t4=time.perf_counter()
    r=ctx.fetch(); t5=time.perf_counter()
    ctx.close(); t6=time.perf_counter()
    print(f"create {1e3*(t1-t0):.1f} ms, first transport {1e3*(t2-t1):.1f}, first finalize {1e3*(t3-t2):.1f}, 19 gens {1e3*(t4-t3):.1f}, fetch {1e3*(t5-t4):.1f}, close {1e3*(t6-t5):.1f}")
    t0=time.perf_counter(); r=nb.monte_carlo(*args,1.0,generations=20,histories=10_000_000,skip=1); t1=time.perf_counter()
    print(f"monte_carlo() {1e3*(t1-t0):.1f} ms, device {1e3*r.seconds_device:.1f} ms")
