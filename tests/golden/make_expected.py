"""Golden OUTPUT vectors for BASELINE configs 1-3, produced by the CPU oracle (oracle/npdr_oracle.c)
on the committed input fixtures.  Test infrastructure.  Regenerate with

    python tests/golden/make_expected.py

NOTE: these are outputs of this repo's restatement of the reference algorithm, not of R (R is not
installable here; see the oracle header).  They pin the oracle against regressions and give the GPU
tests a fixed target; M, the top hits and the IRLS iteration counts agree with the survey-time
numpy restatement recorded in SURVEY.md 8c.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from conftest import load_case, oracle_run  # noqa: E402


def main():
    for name in ("cfg1", "cfg2", "cfg3"):
        case = load_case(name)
        r = oracle_run(case)
        D = r["D"]
        np.savez_compressed(os.path.join(HERE, f"expected_{name}.npz"), Ri=r["Ri"], NN=r["NN"],
                            radius=(r["radius"] if r["radius"] is not None else np.zeros(0)), stats=r["stats"],
                            n_unique=r["n_unique"], D_row0=D[:, 0].copy(), D_sum=float(D.sum()),
                            D_diag_max=float(np.abs(np.diag(D)).max()))
        top = np.argsort(-r["stats"][:, 1])[:3]
        print(name, "M =", r["Ri"].size, "unique =", r["n_unique"], "top:",
              [(case["names"][i], round(float(r["stats"][i, 1]), 3)) for i in top])


if __name__ == "__main__":
    main()
