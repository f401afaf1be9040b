"""Diagnostic: wall time of the bani2v2o8 statistical workload on the GPU."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
from helpers import bani_honeycomb
from sse_b200.walkers import DeviceModel, Walkers
from sse_b200.mc import default_capacity
L = int(sys.argv[1]) if len(sys.argv) > 1 else 10
Ts = np.linspace(0.05, 4, 20)
if L == 20: Ts = Ts[1:]
model = bani_honeycomb(L); dm = DeviceModel(model)
rep = 32
cap, ncap = default_capacity(dm.sse_data, float(Ts.min()))
print("L", L, "cap", cap, flush=True)
gw = Walkers(dm, np.repeat(Ts, rep), m_capacity=cap, n_capacity=ncap, seed=1)
t = time.time(); gw.init(); print("init %.2fs" % (time.time() - t), flush=True)
for i in range(6):
    t = time.time(); gw.sweep(50, thermalized=False); c = gw.fetch_counters(reset=True)
    print("therm 50 sweeps: %.2fs visits/walker-sweep %.0f  mean n %.0f" % (time.time() - t, c["visits"] / c["sweeps"], c["sum_n"] / c["sweeps"]), flush=True)
for i in range(3):
    t = time.time(); gw.sweep(100, thermalized=True, measure=True); c = gw.fetch_counters(reset=True)
    print("meas 100 sweeps: %.2fs visits/walker-sweep %.0f shares %s" % (time.time() - t, c["visits"] / c["sweeps"],
          [round(c[k] / (c["cycles_diag_build"] + c["cycles_worm"] + c["cycles_commit_measure"]), 3) for k in ("cycles_diag_build", "cycles_worm", "cycles_commit_measure")]), flush=True)
n = gw.num_operators().reshape(len(Ts), rep).mean(1)
print("n(T):", np.round(n).astype(int).tolist())
