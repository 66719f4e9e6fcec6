"""deepsc-gan_b200: sm_100a implementation of the DeepSC-GAN transmit path
(encode -> channel(+attack) -> decode -> BLEU) behind the reference's own module surface.

Import as ``deepsc_gan_b200`` (the repository-root shim registers this directory under that name,
because the directory name carries a hyphen).  Layout mirrors the reference:
``models/`` (modules, transceiver, gan), ``utlis/`` (eval, tools, parameters), ``dataset/``;
``csrc/`` holds the CUDA kernels and the C ABI, ``_lib`` the ctypes binding, ``engine``/``sweep`` the
batched greedy decode and the SNR sweep.
"""
__version__ = "0.1.0"
