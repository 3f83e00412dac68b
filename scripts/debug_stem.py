"""Run the stem kernel variants in separate processes (a trap poisons the CUDA context)."""
import subprocess, sys, os
CODE = r'''
import sys, numpy as np, torch
sys.path.insert(0, ".")
import bnn_b200
from bnn_b200 import functional as BF, native
from oracle import c_oracle as co
h, w, flags = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
rng = np.random.default_rng(5)
x = rng.standard_normal((2, 3, h, w)).astype(np.float32)
wt = (rng.standard_normal((64, 3, 7, 7)) * 0.1).astype(np.float32)
g, hh = (0.5 + rng.random(64)).astype(np.float32), (rng.standard_normal(64) * 0.3).astype(np.float32)
d = lambda a: torch.from_numpy(a).cuda()
w_t = BF.stem_weight_layout(d(wt))
out, bits = BF.stem(d(x), w_t, (d(g), d(hh)), flags=flags)
torch.cuda.synchronize()
want_out, want_bits = co.stem(x, wt, g, hh)
print("OK exact_out", np.array_equal(out.permute(0, 2, 3, 1).cpu().numpy(), want_out), "exact_bits",
      np.array_equal(bits.bits.cpu().numpy().view(np.uint32), want_bits))
'''
for h, w, flags in ((64, 64, 1), (64, 64, 0), (224, 224, 0), (64, 64, 8), (64, 64, 4 | 8), (64, 64, 4)):
    r = subprocess.run([sys.executable, "-c", CODE, str(h), str(w), str(flags)], capture_output=True, text=True, timeout=300)
    tail = (r.stdout.strip().splitlines() or [""])[-1] if r.returncode == 0 else (r.stderr.strip().splitlines() or ["?"])[-1][:160]
    print(f"h={h} w={w} flags={flags}: rc={r.returncode} {tail}", flush=True)
