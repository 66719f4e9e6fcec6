"""CPU oracle for the DeepSC-GAN transmit path.  TEST INFRASTRUCTURE ONLY.

This package is a literal PyTorch-CPU restatement of the reference's
TensorFlow/Keras math (DeepSC-GAN/models/*.py, utlis/eval.py, utlis/tools.py).
It is the checker for the CUDA path, never the product:

  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
    ``cpu_baseline`` / ``--impl reference`` legs may import it;
  * nothing under ``deepsc-gan_b200/`` imports it, and the product path raises
    when the CUDA extension is missing instead of falling back here.

PARITY UNPINNED: the reference cannot be imported in this environment
(TensorFlow, nltk and w3lib are not installed; utlis/*.py import symbols that
do not exist; all trained weights are missing) and it ships no tests, golden
vectors or seeds.  The only structural pin is the checkpoint-index parameter
inventory (tests/golden/ckpt_inventory.json, produced by
tests/golden/make_ckpt_inventory.py from DeepSC-GAN/checkpoint/**/ckpt-9.index).
Every TF semantic restated here (Dense kernel layout [in,out], LayerNorm with
biased variance and eps=1e-6, additive -1e9 mask, tf.roll direction, first-max
argmax, Frobenius tf.norm) is written from the TF API contract.
"""
