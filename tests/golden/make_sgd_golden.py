"""Golden vectors for the optimizer step, generated with torch.optim.SGD itself --
the third-party code the reference's run_one_epoch calls (models/yolov2.py:7 `from torch.optim import SGD`,
:1253-1272) -- executing the reference's statement sequence on the CPU:

    opt = SGD(params, lr=..., momentum=0.9, weight_decay=5e-4)   # a NEW optimizer in every iteration
    opt.zero_grad(); loss.backward(); opt.step()

and, for the explicit persistent-momentum mode, ONE optimizer stepped three times.  The loss is a fixed
random function of the parameters so that the gradients depend on them.  (Kept apart from make_golden.py:
that script's import shims for the reference's model files get in the way of torch.optim's lazy imports.)

    python tests/golden/make_sgd_golden.py      ->  tests/golden/sgd_step.npz
"""
import os

import numpy as np
import torch
from torch.optim import SGD

HERE = os.path.dirname(os.path.abspath(__file__))
SHAPES = [(8, 3, 3, 3), (8,), (1,), (7, 5), (16385,), (3, 1, 1, 1)]  # (16385: one element past a chunk of the kernel)
LRS = [1e-3 / (10 ** (1 - 0.25)), 1e-3 / (10 ** (1 - 0.5)), 1e-3]  # the epoch-1 warm-up of :1255, then the plain lr


def loss_of(ps, coef):
    return sum(((p * c).sin() * p).sum() for p, c in zip(ps, coef))


def main(seed=401):
    torch.set_num_threads(1)
    gen = torch.Generator().manual_seed(seed)
    p0 = [torch.randn(s, generator=gen) for s in SHAPES]
    coef = [torch.randn(s, generator=gen) for s in SHAPES]
    out = {"n_tensors": np.int64(len(SHAPES)), "lrs": np.array(LRS), "momentum": 0.9, "weight_decay": 5e-4}
    for i, t in enumerate(p0):
        out["p0_%d" % i] = t.numpy().copy()
    # (a) the reference: a new optimizer per iteration
    ps = [torch.nn.Parameter(t.clone()) for t in p0]
    for it, lr in enumerate(LRS):
        opt = SGD(ps, lr=lr, momentum=0.9, weight_decay=5e-4)
        opt.zero_grad()
        loss_of(ps, coef).backward()
        for i, p in enumerate(ps):
            out["fresh_g%d_%d" % (it, i)] = p.grad.numpy().copy()
        opt.step()
        if it == len(LRS) - 1:  # (the parameters after the last iteration: the earlier ones are implied)
            for i, p in enumerate(ps):
                out["fresh_p%d_%d" % (it, i)] = p.detach().numpy().copy()
    # (b) persistent momentum: one optimizer
    ps = [torch.nn.Parameter(t.clone()) for t in p0]
    opt = SGD(ps, lr=LRS[2], momentum=0.9, weight_decay=5e-4)
    for it in range(3):
        opt.zero_grad()
        loss_of(ps, coef).backward()
        for i, p in enumerate(ps):
            out["pers_g%d_%d" % (it, i)] = p.grad.numpy().copy()
        opt.step()
        if it == 2:
            for i, p in enumerate(ps):
                out["pers_p%d_%d" % (it, i)] = p.detach().numpy().copy()
    np.savez_compressed(os.path.join(HERE, "sgd_step.npz"), **out)
    print("sgd_step.npz written")


if __name__ == "__main__":
    main()
