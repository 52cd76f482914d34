#!/usr/bin/env python3
"""tests/golden/make_goto_state.py -- finds a start state shortly before the multi-ligand alignment takes its `goto lable4` back
edge (main.cpp:1628 -> 1438) and a lay-down (main.cpp:1141) in the same 100 steps, so that a short GPU lockstep window is known to
exercise both. The reference-evolved hot state hot200_step40000.npz is continued by the oracle (bit-equal to the reference,
tests/test_oracle_golden.py) in keyed mode with seed 2024; the state at the start of the first window in which both counters
advance is written to hot200_step180900_goto.npz. The GPU test must use the same seed and regime to meet the same events."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
sys.path.insert(0, os.path.join(HERE, ".."))
import pyoracle  # noqa: E402
from common import apply_regime, load_golden_state  # noqa: E402

g = load_golden_state(os.path.join(HERE, "hot200_step40000.npz"))
o = pyoracle.Oracle(apply_regime(pyoracle.default_params(box=tuple(g["params"]["box"]), n_receptor=150, n_ligand=50, use_grid=1, stream_mode=1, seed=2024), "hot"))
o.set_state(g["R"], g["status"], g["res_nei"], step_done=g["step"], max_complex=g["max_complex"])
prev = o.events().copy()
for w in range(5000):
    st, c = o.get_state(), o.counts()
    o.step(100)
    ev = o.events()
    if ev[9] > prev[9] and ev[8] > prev[8]:
        params = dict(g["params"]); params["keyed_seed"] = 2024
        params["note"] = "oracle-continued start state before a lay-down and a taken goto lable4 (see make_goto_state.py)"
        np.savez_compressed(os.path.join(HERE, "hot200_step%d_goto.npz" % c["step"]), R=st[0], status=st[1], res_nei=st[2], step=c["step"],
                            max_complex=c["max_complex"], params=json.dumps(params))
        print("saved state of step", c["step"])
        break
    prev = ev.copy()
