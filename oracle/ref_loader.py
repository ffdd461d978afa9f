"""Loads the UNMODIFIED reference UNet package (avgen/models/unets) under the private top-level name
`asva_ref_unets`, against oracle/diffusers_shim.  TEST INFRASTRUCTURE ONLY.

The reference tree is looked up at $ASVA_REFERENCE_ROOT, /root/reference, then <repo>/oracle/_ref (a git-ignored
copy that oracle/stage_reference.py makes so the reference's own code can be timed on the GPU box, where
/root/reference does not exist).  All intra-package imports in the reference are relative, so loading it under
another name keeps it from colliding with this repo's own `avgen` boundary package."""
import importlib
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM = os.path.join(_HERE, "diffusers_shim")
_NAME = "asva_ref_unets"


def reference_root():
    cands = [os.environ.get("ASVA_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "avgen", "models", "unets", "__init__.py")):
            return c
    return None


def available() -> bool:
    return reference_root() is not None


def load_reference_unets():
    """Returns the reference module `avgen.models.unets` (exposes AudioUNet3DConditionModel)."""
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    root = reference_root()
    if root is None:
        raise RuntimeError("reference tree not found (set ASVA_REFERENCE_ROOT or run oracle/stage_reference.py)")
    if "diffusers" in sys.modules and not getattr(sys.modules["diffusers"], "__version__", "").endswith("shim"):
        raise RuntimeError("a real diffusers is already imported; the oracle expects the 0.29.2 shim")
    if _SHIM not in sys.path:
        sys.path.insert(0, _SHIM)
    pkg_dir = os.path.join(root, "avgen", "models", "unets")
    spec = importlib.util.spec_from_file_location(_NAME, os.path.join(pkg_dir, "__init__.py"),
                                                  submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod


def build_reference_unet(config: dict):
    """Constructs the reference AudioUNet3DConditionModel (fp32, eval) from constructor kwargs."""
    mod = load_reference_unets()
    model = mod.AudioUNet3DConditionModel(**config)
    model.eval()
    return model
