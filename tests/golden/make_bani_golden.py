#!/usr/bin/env python
"""Extracts the statistical-parity fixture for BASELINE config 4 from the reference's committed result file
/root/reference/docs/src/bani2v2o8.results.json (StochasticSeriesExpansion 0.1.0 / Carlo 0.2.2, S=1 honeycomb,
Dz/J = 0.04556/8.07, L = 10 and 20, 20 temperatures, 80 000 sweeps per task).  Run in the build container
(the reference tree does not exist on the GPU box); writes tests/golden/bani2v2o8_golden.json."""
import json
import os

SRC = "/root/reference/docs/src/bani2v2o8.results.json"
OBS = ["Energy", "SpecificHeat", "Mag", "AbsMag", "Mag2", "Mag4", "MagChi", "BinderRatio", "OperatorCount",
       "WormLengthFraction", "Sign", "_ll_sweep_time", "_ll_measure_time"]


def scalar(x):
    return x[0] if isinstance(x, list) else x


def main():
    tasks = json.load(open(SRC))
    out = []
    for t in tasks:
        p = t["parameters"]
        rec = {"task": os.path.basename(t["task"]), "L": p["lattice"]["size"][0], "T": p["T"], "S": p["S"], "J": p["J"],
               "Dz": p["Dz"], "sweeps": p["sweeps"], "thermalization": p["thermalization"], "binsize": p["binsize"]}
        for o in OBS:
            if o in t["results"]:
                rec[o] = [scalar(t["results"][o]["mean"]), scalar(t["results"][o]["error"])]
        out.append(rec)
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bani2v2o8_golden.json")
    json.dump({"source": SRC, "version": tasks[0].get("version"), "tasks": out}, open(dst, "w"), indent=0)
    print("wrote", dst, len(out), "tasks")


if __name__ == "__main__":
    main()
