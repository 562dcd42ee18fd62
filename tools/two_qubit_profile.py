import os, sys, time, cProfile, pstats, io
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from qucumber_b200 import ComplexWaveFunction, DensityMatrix
from qucumber_b200._lib import load
L = load()
rng = np.random.default_rng(3)
for name, make, n_samples in (("complex", lambda: ComplexWaveFunction(2, 2, gpu=True), 500), ("density", lambda: DensityMatrix(2, 2, 2, gpu=True), 900)):
    bases_all = ["XZ", "ZX", "YZ", "ZY", "ZZ"] if name == "complex" else ["XX", "XY", "XZ", "YX", "YY", "YZ", "ZX", "ZY", "ZZ"]
    bases = np.array([list(bases_all[i % len(bases_all)]) for i in range(n_samples)])
    data = torch.tensor((rng.random((n_samples, 2)) < 0.5).astype(np.float64))
    torch.manual_seed(1)
    nn = make()
    kw = dict(pos_batch_size=100, neg_batch_size=100, k=10, lr=0.1, input_bases=bases)
    nn.fit(data, epochs=2, **kw); torch.cuda.synchronize()
    l0 = L.qcb_launch_count(); t0 = time.perf_counter()
    nn.fit(data, epochs=10, **kw)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{name}: host {(t1-t0)*100:.2f} ms/epoch, with sync {(t2-t0)*100:.2f}; launches/epoch {(L.qcb_launch_count()-l0)/10}")
    pr = cProfile.Profile(); pr.enable(); nn.fit(data, epochs=10, **kw); torch.cuda.synchronize(); pr.disable()
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(30); print(s.getvalue()[:7000])
