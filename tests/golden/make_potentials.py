"""Generates tests/golden/potentials.json: the order-0 result (ga_workspace::assembly(0), assembled_potential()) of the
UNMODIFIED reference for the fixtures of make_golden.py -- same driver arguments, hence the same mesh, state and constants --
with the potential whose first variation is the fixture's order-1 form (oracle/ref_driver.cc mode=potential).

    make -C oracle && python tests/golden/make_potentials.py
"""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import CASES, DRV  # noqa: E402


def main():
    out = {}
    for name, args in CASES.items():
        if "family2" in args or "coef=fem" in args:  # one term per potential; constant coefficients
            continue
        r = subprocess.run([DRV] + args.split() + ["mode=potential"], capture_output=True, text=True)
        if r.returncode != 0:
            print("skipped", name, r.stderr[-200:])
            continue
        d = json.loads(r.stdout.strip().splitlines()[-1])
        out[name] = {"potential": d["potential"], "expr": d["potential_expr"]}
        print(name, d["potential_expr"], d["potential"])
    with open(os.path.join(HERE, "potentials.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
