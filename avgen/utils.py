"""avgen.utils — only what the inference scripts import (/root/reference/avgen/utils.py:13-26)."""


def freeze_model(model):
    for p in model.parameters():
        p.requires_grad_(False)
    return model


def freeze_and_make_eval(model):
    """Stops gradients and switches to eval mode (scripts/animation_demo.py:78 calls it on the audio encoder)."""
    freeze_model(model)
    model.eval()
    return model


def get_model_size(model) -> int:
    return sum(p.numel() for p in model.parameters())
