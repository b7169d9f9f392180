"""consistencytta_b200 — B200 (sm_100a) native implementation of ConsistencyTTA's single-step generation hot path
(guided UNet -> AudioLDM VAE decode -> HiFi-GAN) behind the reference's module API.  See DESIGN.md."""
from .audio_io import save_clips, wav_bytes_pcm16, write_wav_pcm16  # noqa: F401
from .pipeline import AudioLCM, AudioLCM_FTVAE, ConsistencyTTA, SingleStepEngine, build_random_init_models  # noqa: F401
from .scheduler import DDIMScheduler, HeunDiscreteScheduler  # noqa: F401
from .unet import UNet2DConditionGuidedModel, UNet2DConditionOutput  # noqa: F401
from .vae import AutoencoderKL, Decoder, Generator  # noqa: F401

__version__ = "0.1.0"
