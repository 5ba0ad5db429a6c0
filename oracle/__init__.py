"""CPU oracle for the NRMS-family hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package.  The product
(``ebnerd-benchmark_b200/``) never does: it fails loudly when the CUDA
extension is missing.

PARITY UNPINNED: the reference (ebanalyse/ebnerd-benchmark) holds no golden
vector, known-answer test or fixture for the model arithmetic (its tests only
check dataloader batch structure) and TensorFlow is not
installable in the build container, so the restatement below is a line-by-line
numpy restatement of ``src/ebrec/models/newsrec/{layers,nrms,nrms_docvec,naml}.py``
plus the published Keras semantics (Adam, categorical cross-entropy, Dropout,
BatchNormalization, GlorotUniform).  What *is* pinned: the gather-index helpers
and metric known answers quoted in the reference docstrings (tests/golden/).
"""
