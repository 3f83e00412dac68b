"""A geometry whose full-row tile cannot fit shared memory (19 input chunks x 5x5 kernel): runs on narrower tiles now.
Bit-exact integer dots against the oracle."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bnn_b200  # noqa: E402
from bnn_b200 import functional as BF  # noqa: E402
from oracle import c_oracle as co  # noqa: E402

rng = np.random.default_rng(3)
for (cin, cout, k, s, pad, h, w) in ((1153, 32, 5, 2, 2, 5, 57), (1153, 70, 5, 1, 2, 6, 300)):
    x = np.maximum(rng.standard_normal((1, cin, h, w)), 0).astype(np.float32)
    wt = (rng.standard_normal((cout, cin, k, k)) * 0.05).astype(np.float32)
    g = co.geom(1, cin, h, w, cout, k, k, (s, s), (pad, pad), (1, 1))
    wb, alpha, nz = co.pack_weight(wt, True, True)
    assert nz == 0
    want = co.bconv2d(co.pack_act(x), wb, None, None, None, g)
    act = BF.pack_activations(torch.from_numpy(x).cuda())
    wts = BF.pack_weights(torch.from_numpy(wt).cuda(), True, True)
    got = BF.bconv2d(act, wts, None, None, (s, s), (pad, pad), (1, 1), use_alpha=False).cpu().numpy()
    assert np.array_equal(got, want), (cin, cout, k, s, np.abs(got - want).max())
    print("narrow-tile geometry ok", (cin, cout, k, s, pad, h, w), "plan", bnn_b200.native.conv_plan(
        bnn_b200.native.ConvGeom(1, cin, h, w, cout, k, k, s, s, pad, pad, 1, 1), 0, 148))
