"""CPU oracle for the SAGAN generator/discriminator hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`self-attention-gan_b200/`) imports this directory; only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may use it, and there only as the checker / the timed CPU baseline.

What it is: a plain numpy / torch-CPU restatement (fp64 or fp32, un-fused, op by
op like the reference's TF graph) of the reference algorithm:

  * `oracle.sn`         <- /root/reference/layers.py:4-68   (l2normalize, power iteration, W/sigma)
  * `oracle.attention`  <- /root/reference/layers.py:71-120 (Attention_Layer)
  * `oracle.nets`       <- /root/reference/sagan/models/generator.py:7-37,
                           /root/reference/sagan/models/discriminator.py:7-36
  * `oracle.train`      <- /root/reference/sagan/main.py:21-27,111-120,171-229
  * `oracle.resnets`    <- /root/reference/models/generator.py:6-43, models/discriminator.py:6-57
  * `oracle.weightnorm` <- /root/reference/sagan/layers.py:6-211, sagan/dataset.py:27-40

PARITY: PINNED TO THE REFERENCE'S OWN CODE FOR THE FORWARD PATH, THE STEP SCHEDULE
AND THE LOSS SCALING; UNPINNED FOR THE GRADIENTS AND THE OPTIMISER ARITHMETIC.
The reference ships no golden vectors, known-answer tests or published numbers
for this path (its tests assert output shapes only, test/test_generator.py:26,
test/test_discriminator.py:28) and its arithmetic primitives live in an absent,
un-pinned third-party dependency (`tensorflow`, TF 2.0-era API; not installable
in this image, no network).  What CAN run here is the reference's Python itself:
`tests/golden/make_reference_vectors.py` imports /root/reference/layers.py,
sagan/models/generator.py, sagan/models/discriminator.py and the hinge losses of
sagan/main.py:21-27 UNMODIFIED, with `tests/golden/tfshim/tensorflow` (a float64
numpy stand-in for the few tf / Keras symbols they use) on the import path, and
freezes what they compute in `tests/golden/reference_layers.npz`:
l2normalize; SpectralNormalization._make_param / update_uv (u, v, sigma,
W / sigma over two calls, Keras Dense / Conv2D / Conv2DTranspose kernel layouts,
Ip 1-3, factor); Attention_Layer.build / call; hinge_loss_d / hinge_loss_g;
get_generator / get_discriminator (patch head and projection head) layer by
layer; the legacy residual builders models/generator.py, models/discriminator.py
(attention at C = 8 inside); the weight-normalisation wrapper of sagan/layers.py
(data-dependent and norm initialisation, two calls); the record reader of
sagan/dataset.py:12-40 (float32 decode, bit for bit); Trainer.train_step /
distributed_train_step of sagan/main.py:171-236 on stub models (call schedule,
differentiated scalars, reported losses).  `tests/test_reference_vectors.py` holds this oracle to those numbers
at 1e-12 and the CUDA kernels at the north_star tolerances.

Still unpinned (no reference code to execute, or TF itself needed): TensorFlow's
own kernels under those calls (matmul / conv / softmax semantics come from the
stand-in, per their documentation); the gradients (tf.GradientTape; the oracle's
analytic gradients are checked against torch autograd and finite differences
instead); Keras Adam / ExponentialDecay (sagan/main.py:111-120: library code,
restated from its documentation); attention at
C > 8 in the literal code (its MaxPool2D(2, 1) + raw reshape is ill-formed, see
make_reference_vectors.py).  The readings chosen where the literal reference
code is ill-formed are listed in DESIGN.md ("Oracle readings").
"""
