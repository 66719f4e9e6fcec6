"""Freeze oracle outputs for the seeded parity cases (tests/_cases.py) into tests/golden/*.npz.

    python tests/golden/make_goldens.py

PARITY UNPINNED: these goldens come from the oracle (oracle/deepsc_oracle.py), not from the reference,
which cannot run here.  They pin the oracle against drift and give the GPU tests a
fixture that does not need the oracle's CPU time.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import _cases  # noqa: E402


def main():
    for kind, channel in [("Transeiver_Star", "AWGN"), ("Transeiver_Star", "Rayleigh"), ("Transeiver", "AWGN"),
                          ("Transeiver_star", "AWGN"), ("Transeiver_GAN", "AWGN")]:
        c = _cases.oracle_case(kind, channel)
        keep = {k: v for k, v in c.items()}
        keep["symbols"] = keep["symbols"][:8]
        keep["received"] = keep["received"][:8]
        path = _cases.golden_path(kind, channel)
        np.savez_compressed(path, **keep)
        print(kind, channel, {k: getattr(v, "shape", None) for k, v in keep.items()}, os.path.getsize(path))


if __name__ == "__main__":
    main()
