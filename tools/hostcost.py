import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch, swarm_b200
for N, E in ((32, 64), (32, 1024), (8, 256)):
    eng = swarm_b200.SwarmEngine(E, {"num_drones": N, "num_obstacles": 8 if N == 32 else 4}, device="cuda:0")
    eng.seed(np.arange(E, dtype=np.uint64)); eng.reset()
    a = torch.rand((E, N, 3), device="cuda:0") * 2 - 1
    for _ in range(50): eng.step(a)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(2000): eng.step(a)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"N={N} E={E}: host enqueue {(t1-t0)/2000*1e6:.1f} us/step, total {(t2-t0)/2000*1e6:.1f} us/step")
