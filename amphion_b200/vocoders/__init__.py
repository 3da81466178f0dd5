"""Native GAN vocoders behind the reference's registry/dispatch surface
(models/vocoders/vocoder_inference.py:39-75)."""
from .apnet import APNet
from .bigvgan import BigVGAN
from .hifigan import HiFiGAN, HiFiGAN_vits
from .nsfhifigan import NSFHiFiGAN
from .gan_vocoder_inference import synthesis_audios, synthesize_to_files, vocoder_inference

# same shape as the reference's registries: generator name -> class / functions
_vocoders = {"hifigan": HiFiGAN, "bigvgan": BigVGAN, "nsfhifigan": NSFHiFiGAN, "apnet": APNet}
_vocoder_forward_funcs = {k: vocoder_inference for k in _vocoders}
_vocoder_infer_funcs = {k: synthesis_audios for k in _vocoders}

__all__ = ["HiFiGAN", "HiFiGAN_vits", "BigVGAN", "NSFHiFiGAN", "APNet", "vocoder_inference", "synthesis_audios", "synthesize_to_files", "_vocoders",
           "_vocoder_forward_funcs", "_vocoder_infer_funcs"]
