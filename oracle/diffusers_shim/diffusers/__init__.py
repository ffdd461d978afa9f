"""ORACLE / TEST INFRASTRUCTURE ONLY — never imported by the product package (asva_b200/, avgen/).

A ~250-line restatement of the handful of diffusers==0.29.2 classes that the reference's UNet files import
(/root/reference/requirements.txt:2 pins the version; diffusers is not installed in this image and cannot be:
no network).  With this directory and the reference checkout on sys.path, the UNMODIFIED reference files under
avgen/models/unets import and run on CPU; that run is what pins the clean-room oracle (oracle/unet_ref.py) and
what produces tests/golden/.  Semantics restated from the published diffusers 0.29.2 sources:
  models/attention_processor.py  Attention, AttnProcessor2_0
  models/attention.py            FeedForward, GEGLU
  models/embeddings.py           get_timestep_embedding, Timesteps, TimestepEmbedding
  configuration_utils.py         ConfigMixin, register_to_config
  models/modeling_utils.py       ModelMixin (.device / .dtype)
  utils/outputs.py               BaseOutput
"""
__version__ = "0.29.2-shim"
