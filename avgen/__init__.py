"""Drop-in `avgen` package for ASVA's denoising hot path.

This package provides ONLY the hot path (avgen.models.unets, avgen.pipelines.pipeline_audio_cond_animation,
avgen.utils) on the B200 engine in asva_b200/.  Everything else the reference's scripts import
(avgen.data, avgen.models.audio_encoders, avgen.evaluations, ... - data loading, ImageBind, metrics; out of scope
here) keeps coming from the reference checkout: put it on sys.path AFTER this repo and the line below splices its
`avgen/` directory into this package's search path, so `scripts/animation_demo.py` / `animation_gen.py` run
unmodified."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
