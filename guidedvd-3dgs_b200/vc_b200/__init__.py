"""B200-native building blocks of the ViewCrafter denoiser (U-Net) and DDIM sampler."""
