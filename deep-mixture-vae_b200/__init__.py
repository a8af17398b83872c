"""dmvae_b200 - B200-native (sm_100a) training step of the Deep Mixture VAE family.

Drop-in for the hot path of ffs97/deep-mixture-vae (code/base_models.py, code/priors.py, code/models.py and the
train.py flags): hand-written CUDA behind a C ABI (include/dmvae_b200.h), PyTorch only for device memory,
streams and torch.distributed.  No Triton, no CPU fallback.
"""
__version__ = "0.1.0"
