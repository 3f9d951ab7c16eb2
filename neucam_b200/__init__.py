"""neucam_b200 -- B200-native (sm_100a) implementation of NeuCam's training hot path.

The package mirrors the reference's `modules.py` nn.Module surface (SingleBVPNet / FCBlock /
BatchLinear / Sine, SimpleMLP, TonemapNet, LearnExps, LearnFocal) and the per-step pipeline of its
`training.py`, executed by hand-written CUDA kernels behind a C ABI (include/neucam_b200.h).
"""
__version__ = "0.1.0"
