"""TEST INFRASTRUCTURE: ptflops stand-in (utils/utils.py:6; only `print_model_stats`, which no script calls)."""


def get_model_complexity_info(model, input_res, as_strings=True, print_per_layer_stat=False, **kw):
    params = sum(p.numel() for p in model.parameters())
    return ("0 GMac", f"{params / 1e6:.2f} M") if as_strings else (0.0, params)
