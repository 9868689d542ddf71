"""GPU forward vs the bf16-emulating oracle (oracle.forward_bf16_emulated): per-module tap statistics and logits.
usage: python tools/emu_check.py [batch] [mask_mode]"""
import sys

import numpy as np

sys.path.insert(0, ".")
import ishara_b200 as ib
from oracle import ishara_oracle as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
mode = sys.argv[2] if len(sys.argv) > 2 else "dropped"
cfg = O.Config()
params = O.init_params(cfg, seed=42)
x = O.make_inputs(cfg, B, seed=7, ragged=(mode == "propagated"))
m = ib.get_model(mask_mode=mode).load_weights(params)
taps_e, taps_64 = {}, {}
emu = O.forward_bf16_emulated(params, x, cfg, mask_mode=mode, taps=taps_e)
ref = O.forward(params, x, cfg, "float64", mask_mode=mode, taps=taps_64)
got = m(x)
taps_g = m.debug_activations(x, list(taps_e))


def stats(a, b):
    d = np.abs(a.astype(np.float64) - b.astype(np.float64))
    return d.max() / np.abs(b).max(), (a == b).mean()


for k in taps_e:
    if k in taps_g:
        r_e, same = stats(taps_g[k], taps_e[k])
        r_64, _ = stats(taps_g[k], taps_64[k])
        print(f"{k:24s} vs emu rel={r_e:.3g} identical={same:.4f} | vs fp64 rel={r_64:.3g}")
for name, r in (("emu", emu), ("fp64", ref)):
    err = np.abs(got - r).max() / np.abs(r).max()
    agree = (got.argmax(-1) == r.argmax(-1)).mean()
    top2 = np.sort(r, -1)
    margin = top2[..., -1] - top2[..., -2]
    dis = got.argmax(-1) != r.argmax(-1)
    print(f"logits vs {name}: rel={err:.4g} agree={agree:.5f} rms={np.sqrt(np.mean((got - r) ** 2)) / np.abs(r).max():.4g}"
          f" max margin at a disagreeing frame={(margin[dis].max() if dis.any() else 0):.4g} (scale {np.abs(r).max():.3g})")
