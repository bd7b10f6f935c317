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

PARITY UNPINNED: the reference ships no golden vectors, known-answer tests or
published numbers for this path (its tests assert output shapes only,
test/test_generator.py:26, test/test_discriminator.py:28) and its arithmetic
lives in an absent, un-pinned third-party dependency (`tensorflow`, TF 2.0-era
API; not installed in this image, no network).  The restatement is therefore
anchored on the reference's own call sites and on self-consistency checks
(finite differences, torch-autograd vs analytic gradients, fp32 vs fp64); the
readings chosen where the literal reference code is ill-formed are listed in
DESIGN.md ("Oracle readings").
"""
