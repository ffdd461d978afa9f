from .audio_cond_unet_3d_condition import AudioUNet3DConditionModel, UNet3DConditionOutput  # noqa: F401
