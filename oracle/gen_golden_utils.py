"""Golden vectors for the small host-side helpers of utils.py that the hot path's callers use (reshape_tensor :2770,
reshape_encoder_output_average :2776, average_over_W_H :1996, variance_explained :2012, get_top_k_samples :1445,
CustomCrossEntropyLoss :99): runs the REAL reference functions (ref_import.py) on seeded inputs and freezes inputs and
outputs in tests/golden/utils_small.npz.  TEST INFRASTRUCTURE ONLY; needs /root/reference, so it runs in the build
container and its output is committed."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402


def main():
    R = ref_import.load()
    g = torch.Generator().manual_seed(77)
    out = {}
    x4 = torch.randn(3, 6, 4, 5, generator=g)
    d4 = x4 + 0.1 * torch.randn(3, 6, 4, 5, generator=g)
    x2 = torch.randn(7, 6, generator=g)
    d2 = x2 + 0.1 * torch.randn(7, 6, generator=g)
    out["x4"], out["d4"], out["x2"], out["d2"] = x4, d4, x2, d2
    t, flag = R.reshape_tensor(x4)
    out["reshape4"], out["reshape4_flag"] = t, torch.tensor(int(bool(flag)))
    t, flag = R.reshape_tensor(x2)
    out["reshape2"], out["reshape2_flag"] = t, torch.tensor(int(bool(flag)))
    avg = torch.randn(8, 4, 5, generator=g)                   # [F, H, W]
    out["avg"] = avg
    out["avg_reshaped_b3"] = R.reshape_encoder_output_average(avg, 3)
    a, b = R.average_over_W_H(x4, d4)
    out["avgwh_a"], out["avgwh_b"] = a, b
    a, b = R.average_over_W_H(x2, None)
    out["avgwh2_a"] = a
    assert b is None
    out["var_expl4"] = torch.as_tensor(R.variance_explained(x4, d4))
    out["var_expl2"] = torch.as_tensor(R.variance_explained(x2, d2))
    # get_top_k_samples: running top-k over three batches, largest and smallest
    k, bs, F = 4, 5, 6
    for largest in (True, False):
        state = (torch.empty(0, F), torch.empty(0, F, dtype=torch.long), bs, torch.empty(0, F, dtype=torch.long))
        for batch in (1, 2, 3):
            vals = torch.randn(bs, F, generator=g)
            v, i = torch.topk(vals, k=k, dim=0, largest=largest)
            files = torch.randint(0, 1000, (k, F), generator=g)
            tag = f"topk_{int(largest)}_{batch}"
            out[tag + "_v"], out[tag + "_i"], out[tag + "_f"] = v.clone(), i.clone(), files.clone()
            state = R.get_top_k_samples(state, v, i, files, batch, largest, k)
            out[tag + "_out_v"], out[tag + "_out_i"], out[tag + "_out_f"] = state[0].clone(), state[1].clone(), state[3].clone()
            assert state[2] == bs
    probs = torch.softmax(torch.randn(9, 10, generator=g), dim=1)
    tgt = torch.randint(0, 10, (9,), generator=g)
    out["nll_probs"], out["nll_targets"] = probs, tgt
    out["nll"] = R.CustomCrossEntropyLoss()(probs, tgt)
    path = os.path.join(ROOT, "tests", "golden", "utils_small.npz")
    np.savez_compressed(path, **{k: v.detach().numpy() for k, v in out.items()})
    print("wrote", path, len(out), "arrays")


if __name__ == "__main__":
    main()
