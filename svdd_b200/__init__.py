"""svdd_b200 -- B200-native SVDD decoding engine (drop-in for the decode path of
masa-ue/SVDD: ``Diffusion.controlled_sample*`` behind ``decode.py`` /
``decode_tweedie.py``).  Hand-written sm_100a kernels behind a C ABI
(include/svdd_b200.h); PyTorch is plumbing only.  No CPU fallback."""
__version__ = '0.1.0'
