import os, sys
sys.path.insert(0, "/root/repo")
from odexp_b200 import capi
lib = capi.ModelLib.from_pop("/root/repo/models/kuramoto_lowrank64.pop")
n = 1 << 20
for pad in (0, 32, 224, 1056, 4128, 33*32*31):
    pop = capi.Population(lib, capacity=n + pad)
    pop.init(n, replay=False)
    pop.rhs(0.0); pop.sync()
    best = 1e9
    for rep in range(3):
        pop.event_record(0)
        for _ in range(5):
            pop.rhs(0.0)
        pop.event_record(1)
        best = min(best, pop.event_elapsed_ms(0, 1) / 5)
    print("pad %6d: %.3f ms per RHS evaluation" % (pad, best), flush=True)
    pop.close()
