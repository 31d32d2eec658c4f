"""CPU oracle for the S5 NDNS forward path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``sparsernns_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
as the timed CPU baseline, never as the product.

What it is: a NumPy restatement of the reference algorithm (stevenabreu7/
SparseRNNs) for the three numerics flavours of the S5 NDNS model forward:

* ``fxp``        - fixed-point tensor type + ops      (sparseRNNs/fxparray.py)
* ``quant``      - static fake-quant primitives        (sparseRNNs/utils/quantization.py:352-517, 673-816)
* ``ssm``        - ZOH discretisation, float / static-quant S5 SSM
                   (sparseRNNs/model/ssm.py:37-345, 525-622)
* ``model``      - S5 block + regression model shell   (model/layers.py:180-241, model/seq_model.py:100-115, 439-456)
* ``fxpmodel``   - integer-only model                  (sparseRNNs/fxpmodel.py)
* ``fxpconfig``  - fxp_qconfig construction            (sparseRNNs/fxputils.py:121-142, 155-285, 351-401, 453-786)
* ``ndns``       - STFT / iSTFT / SI-SNR / loss tail      (sparseRNNs/train_helpers.py:16-53, 1381-1412, 1570-1614);
                   PINNED against scipy.signal.stft / istft (tests/test_oracle_ndns.py)
* ``synth``      - recipe-shaped random model, ERK magnitude masks, one-batch
                   calibration

PARITY PIN STATUS: pinned against outputs of the reference's OWN sources, executed in place
from /root/reference under a NumPy-backed stand-in for the part of jax.numpy / jax.lax they use
(tests/golden/jaxshim.py, tests/golden/make_reference_golden.py; vectors committed as
tests/golden/reference_*.npz and checked by tests/test_oracle_vs_reference.py): 84 fixed-point op
cases, the qconfigs, every sow()-named intermediate of FxpRegressionModel.forward, the float and
static-quant SSM.  The ``ndns`` module is pinned against SciPy.  NOT pinned: XLA itself - the last
ulp of its float32 log2 at a compute_best boundary and its fp32 dot accumulation order (the float
flavours carry tolerances for that).  The reference ships no golden vectors of its own (SURVEY.md
section 8c) and cannot be installed here.  Further pins: the reference's
inline property check of the quantiser (quantization.py:900-910) and hand-derived known-answer
cases for every fixed-point primitive (tests/test_oracle_kat.py).
Every function cites the reference file:line it follows.

JAX semantics that NumPy does not give for free and that are enforced here:
int32 payloads that wrap on overflow, arithmetic ``>>``, round-half-to-even,
float32 ``log2``/``ceil`` for data-dependent exponents, lexicographic complex
``maximum`` for ReLU on complex values.
"""
