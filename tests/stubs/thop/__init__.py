"""TEST INFRASTRUCTURE: thop stand-in (utils/utils.py:4-5,116-117).  Like the real `profile` it registers forward hooks on
every leaf module and runs ONE forward of the model; the op count is whatever the hooks saw (the hand-written kernels of
cenet_b200 do not go through nn.Module leaves, so it is 0 there), params are counted from the module."""
import torch


def profile(model, inputs=(), custom_ops=None, verbose=True, **kw):
    seen = [0]
    hooks = []
    for m in model.modules():
        if len(list(m.children())) == 0:
            hooks.append(m.register_forward_hook(lambda mod, i, o: seen.__setitem__(0, seen[0] + 1)))
    was = model.training
    model.eval()
    with torch.no_grad():
        model(*inputs)
    model.train(was)
    for h in hooks:
        h.remove()
    params = sum(p.numel() for p in model.parameters())
    return float(seen[0]), float(params)


def clever_format(nums, fmt="%.2f"):
    out = []
    for n in nums:
        for div, suf in ((1e12, "T"), (1e9, "G"), (1e6, "M"), (1e3, "K")):
            if n >= div:
                out.append((fmt % (n / div)) + suf)
                break
        else:
            out.append(fmt % n)
    return out if len(out) > 1 else out[0]
