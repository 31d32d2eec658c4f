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

PARITY PIN STATUS: **parity unpinned by the reference** (model path; the ``ndns`` module is pinned against SciPy).  The reference ships no
tests directory, golden vectors or fixtures for this path (SURVEY.md section 4,
section 8c) and JAX/Flax are not importable in the build container, so the
reference itself cannot be executed to generate vectors.  The oracle is pinned
instead by (i) the reference's own inline property checks for the quantiser
(quantization.py:900-910) restated in ``tests/test_oracle_kat.py``, and (ii)
hand-derived known-answer cases written from the reference source for every
fixed-point primitive.  Every function cites the reference file:line it follows.

JAX semantics that NumPy does not give for free and that are enforced here:
int32 payloads that wrap on overflow, arithmetic ``>>``, round-half-to-even,
float32 ``log2``/``ceil`` for data-dependent exponents, lexicographic complex
``maximum`` for ReLU on complex values.
"""
